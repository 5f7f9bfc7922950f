/// bed_cuda.d — what the reference's liftBED (source/swiftover/bed.d:72-204) becomes on top
/// of libswiftover_cuda: read all -> swo_lift_bed_text -> write all.  NOT BUILT here (no D
/// toolchain in this image); the C++ host in ../host/ is the executable twin.
module swiftover.bed_cuda;

import std.exception : enforce;
import std.stdio;
import std.string : fromStringz, toStringz;

import swiftover.cuda;

void liftBED(string chainfile, string infile, string outfile, string unmatched)
{
    swo_ctx* ctx;
    enforce(swo_create(&ctx, null, 0) == SWO_OK, "no CUDA device (there is no CPU fallback)");
    scope (exit) swo_destroy(ctx);
    enforce(swo_load_chain(ctx, chainfile.toStringz) == SWO_OK, swo_last_error(ctx).fromStringz);

    File fi = (infile == "-" || infile == "") ? stdin : File(infile, "r");   // bed.d:89-92
    // pinned staging so the GC cannot move the buffer during the call and H2D runs at full rate
    ubyte[] data;
    foreach (chunk; fi.byChunk(1 << 22)) data ~= chunk;
    auto pinned = cast(char*) swo_host_alloc(data.length);
    scope (exit) swo_host_free(pinned);
    pinned[0 .. data.length] = cast(char[]) data[];

    swo_text_out o;
    enforce(swo_lift_bed_text(ctx, pinned, data.length, &o) == SWO_OK, swo_last_error(ctx).fromStringz);

    File fo = (outfile == "-" || outfile == "") ? stdout : File(outfile, "w"); // bed.d:94-97
    fo.rawWrite(o.lifted[0 .. o.lifted_len]);
    if (unmatched) File(unmatched, "w").rawWrite(o.unmatched[0 .. o.unmatched_len]); // bed.d:99-100
}
