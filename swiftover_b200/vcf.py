"""liftVCF(chainfile, genomefile, infile, outfile, unmatched) — mirror of the reference's VCF
driver (source/swiftover/vcf.d:29-176) for uncompressed text VCF and FASTA: read all -> one
swo_lift_vcf_text call (header rewrite + column split on the host, liftDirectly + REF check in
lift_vcf_points_kernel) -> write all."""
from .chain import ChainFile


def liftVCF(chainfile, genomefile, infile, outfile, unmatched, devices=None, filedate=None):
    cf = ChainFile(chainfile, devices=devices)
    try:
        cf.load_genome_fasta(genomefile)  # vcf.d:38
        with open(infile, "rb") as f:
            data = f.read()
        lifted, unm, stats = cf.lift_vcf_text(data, chainfile, genomefile, filedate)
        with open(outfile, "wb") as f:
            f.write(lifted)
        with open(unmatched, "wb") as f:
            f.write(unm)
        return stats
    finally:
        cf.close()
