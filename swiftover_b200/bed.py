"""liftBED(chainfile, infile, outfile, unmatched) — mirror of the reference's BED driver
(source/swiftover/bed.d:72-204): read all -> one swo_lift_bed_text call -> write all."""
import sys

from .chain import ChainFile


def liftBED(chainfile, infile, outfile, unmatched, devices=None):
    cf = ChainFile(chainfile, devices=devices)
    try:
        if infile in ("-", ""):  # bed.d:89-92
            data = sys.stdin.buffer.read()
        else:
            with open(infile, "rb") as f:
                data = f.read()
        lifted, unm, stats = cf.lift_bed_text(data)
        if outfile in ("-", ""):  # bed.d:94-97
            sys.stdout.buffer.write(lifted)
            sys.stdout.buffer.flush()
        else:
            with open(outfile, "wb") as f:
                f.write(lifted)
        if unmatched:  # bed.d:99-100
            with open(unmatched, "wb") as f:
                f.write(unm)
        return stats
    finally:
        cf.close()
