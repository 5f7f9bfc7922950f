/// maf_cuda.d — what the reference's liftMAF (source/swiftover/maf.d:119-319) becomes on top of
/// libswiftover_cuda: chain + genome -> read all -> swo_lift_maf_text -> write all.  The two header lines
/// (maf.d:166-171), the 0 / 1 / many rule (maf.d:192-298), the REF fetch and compare (maf.d:263-265) and the
/// reverse complement of the tumor alleles (maf.d:21-54) are inside the call.
/// NOT BUILT here (no D toolchain in this image); ../host/maf.cpp is the executable twin.
module swiftover.maf_cuda;

import std.exception : enforce;
import std.stdio;
import std.string : fromStringz, toStringz;

import swiftover.cuda;

void liftMAF(string chainfile, string genomefile, string genomebuild, string infile, string outfile,
             string unmatched)
{
    stderr.writeln("Only fields build, chr, start, end, ref/tumor1/tumor2 are updated.");          // maf.d:123
    stderr.writeln("A reference allele change might invalidate other fields.");                    // maf.d:124
    swo_ctx* ctx;
    enforce(swo_create(&ctx, null, 0) == SWO_OK, "no CUDA device (there is no CPU fallback)");
    scope (exit) swo_destroy(ctx);
    enforce(swo_load_chain(ctx, chainfile.toStringz) == SWO_OK, swo_last_error(ctx).fromStringz);   // maf.d:132
    enforce(swo_load_genome_fasta(ctx, genomefile.toStringz) == SWO_OK, swo_last_error(ctx).fromStringz); // maf.d:126

    File fi = (infile == "-" || infile == "") ? stdin : File(infile, "r");   // maf.d:137-140
    ubyte[] data;
    foreach (chunk; fi.byChunk(1 << 22)) data ~= chunk;
    auto pinned = cast(char*) swo_host_alloc(data.length);
    scope (exit) swo_host_free(pinned);
    pinned[0 .. data.length] = cast(char[]) data[];

    swo_text_out o;
    swo_maf_stats st;
    enforce(swo_lift_maf_text(ctx, pinned, data.length, genomebuild.toStringz, &o, &st) == SWO_OK,
            swo_last_error(ctx).fromStringz);

    File fo = (outfile == "-" || outfile == "") ? stdout : File(outfile, "w"); // maf.d:142-145
    fo.rawWrite(o.lifted[0 .. o.lifted_len]);
    if (unmatched) File(unmatched, "w").rawWrite(o.unmatched[0 .. o.unmatched_len]); // maf.d:147-148
    // the closing log lines of maf.d:310-318
    stderr.writefln("Matched %d records (%d multiply, skipped) and %d unmatched", st.n_matched, st.n_multiple, st.n_unmatched);
    stderr.writefln("%d REF allele changes (%d not matching either tumor allele)", st.n_refchg, st.n_neither);
    if (st.n_matched == 0)
        stderr.writeln("Could you have a chromosome name mismatch between chain file and MAF?");
}
