"""liftMAF(chainfile, genomefile, genomebuild, infile, outfile, unmatched) — mirror of the
reference's MAF driver (source/swiftover/maf.d:119-319): read all -> one swo_lift_maf_text call
(column split / join and tumor-allele reverse complement on the host; lift, REF fetch and
compare in lift_maf_records_kernel) -> write all."""
from .chain import ChainFile


def liftMAF(chainfile, genomefile, genomebuild, infile, outfile, unmatched, devices=None):
    cf = ChainFile(chainfile, devices=devices)
    try:
        cf.load_genome_fasta(genomefile)  # maf.d:126
        with open(infile, "rb") as f:
            data = f.read()
        lifted, unm, stats = cf.lift_maf_text(data, genomebuild)
        with open(outfile, "wb") as f:
            f.write(lifted)
        with open(unmatched, "wb") as f:
            f.write(unm)
        return stats
    finally:
        cf.close()
