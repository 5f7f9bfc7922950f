/// vcf_cuda.d — what the reference's liftVCF (source/swiftover/vcf.d:29-169) becomes on top of
/// libswiftover_cuda: chain + genome -> read all -> swo_lift_vcf_text -> write all.  The header rewrite
/// (vcf.d:59-108, contigSort vcf.d:181-215) and the record body (vcf.d:115-169) are inside the call.
/// NOT BUILT here (no D toolchain in this image); ../host/vcf.cpp is the executable twin.
/// Text VCF and plain FASTA only: the library has no htslib (DESIGN.md section 8).
module swiftover.vcf_cuda;

import std.datetime.systime : Clock;
import std.exception : enforce;
import std.format : format;
import std.stdio;
import std.string : fromStringz, toStringz;

import swiftover.cuda;

void liftVCF(string chainfile, string genomefile, string infile, string outfile, string unmatched,
             bool removeTags = false)
{
    enforce(!removeTags, "removeTags is not implemented upstream either (vcf.d:33)");
    swo_ctx* ctx;
    enforce(swo_create(&ctx, null, 0) == SWO_OK, "no CUDA device (there is no CPU fallback)");
    scope (exit) swo_destroy(ctx);
    enforce(swo_load_chain(ctx, chainfile.toStringz) == SWO_OK, swo_last_error(ctx).fromStringz);   // vcf.d:36
    enforce(swo_load_genome_fasta(ctx, genomefile.toStringz) == SWO_OK, swo_last_error(ctx).fromStringz); // vcf.d:38

    File fi = (infile == "-" || infile == "") ? stdin : File(infile, "r");   // vcf.d:44-47
    ubyte[] data;
    foreach (chunk; fi.byChunk(1 << 22)) data ~= chunk;
    auto pinned = cast(char*) swo_host_alloc(data.length);
    scope (exit) swo_host_free(pinned);
    pinned[0 .. data.length] = cast(char[]) data[];

    // ##fileDate of the rewritten header (vcf.d:63-66)
    auto now = Clock.currTime;
    auto filedate = format("%04d%02d%02d", now.year, cast(int) now.month, now.day);

    swo_text_out o;
    enforce(swo_lift_vcf_text(ctx, pinned, data.length, chainfile.toStringz, genomefile.toStringz,
                              filedate.toStringz, &o) == SWO_OK, swo_last_error(ctx).fromStringz);

    File fo = (outfile == "-" || outfile == "") ? stdout : File(outfile, "w"); // vcf.d:49-52
    fo.rawWrite(o.lifted[0 .. o.lifted_len]);
    if (unmatched) File(unmatched, "w").rawWrite(o.unmatched[0 .. o.unmatched_len]); // vcf.d:54-57
}
