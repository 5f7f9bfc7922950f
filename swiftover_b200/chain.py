"""Host-side mirror of the reference's `ChainFile` (source/swiftover/chain.d:485-714)
on top of the C ABI.  Same member names and argument meaning; the work is done by the
sm_100a kernels in libswiftover_cuda.so.  Batched entry points (`lift_intervals`,
`lift_points`, `lift_bed_text`) expose the same operations on whole arrays / buffers.
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from . import _lib as L

#: one lifted piece of an interval: destination contig, ordered start/end (bed.d:165-168),
#: strand byte after the cumulative STRAND_TABLE flip (bed.d:173) or b"" when no strand given
LiftedInterval = namedtuple("LiftedInterval", "contig start end strand")



def _bytes_at(ptr, n):
    """bytes(n) copied from a host address; ctypes.string_at takes a C int size and fails past 2 GiB
    (outputs of heavy fan-out inputs do get that large)."""
    return bytes((C.c_char * n).from_address(ptr)) if n else b""

class SwiftoverError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"swiftover_cuda error {code}: {msg}")
        self.code = code


def _addr(a):
    return a.ctypes.data if a is not None else None


class ChainFile:
    """ChainFile(fn): parse a UCSC chain file into the device-resident block table
    (chain.d:509-597).  `devices` lists the CUDA ordinals to replicate it on."""

    def __init__(self, fn=None, *, text=None, devices=None, inspect_only=False):
        self._lib = L.lib()
        h = C.c_void_p()
        if inspect_only:  # host parser only; every lift fails (tests of the chain parser)
            rc = self._lib.swo_create_inspect(C.byref(h))
        elif devices:
            arr = (C.c_int * len(devices))(*devices)
            rc = self._lib.swo_create(C.byref(h), arr, len(devices))
        else:
            rc = self._lib.swo_create(C.byref(h), None, 0)
        if rc != L.SWO_OK:
            raise SwiftoverError(rc, "swo_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h
        if fn is not None:
            self._ck(self._lib.swo_load_chain(self._h, fn.encode()))
        elif text is not None:
            self._ck(self._lib.swo_load_chain_mem(self._h, text, len(text)))
        else:
            raise ValueError("ChainFile needs a file name or text=")
        lib = self._lib
        #: ChainFile.contigNames (chain.d:500)
        self.contigNames = [lib.swo_qcontig_name(h, i) for i in range(lib.swo_num_qcontigs(h))]
        #: ChainFile.qContigSizes (chain.d:504)
        self.qContigSizes = {lib.swo_qsize_name(h, i): lib.swo_qsize_value(h, i) for i in range(lib.swo_num_qsizes(h))}
        self.sourceContigs = [lib.swo_tcontig_name(h, i) for i in range(lib.swo_num_tcontigs(h))]
        self._tcid = {n: i for i, n in enumerate(self.sourceContigs)}

    def _ck(self, rc):
        if rc != L.SWO_OK:
            raise SwiftoverError(rc, self._lib.swo_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.swo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- table inspection --------------------------------------------------------
    @property
    def handle(self):
        return self._h

    @property
    def num_blocks(self):
        return self._lib.swo_num_blocks(self._h)

    @property
    def is_disjoint(self):
        return bool(self._lib.swo_table_is_disjoint(self._h))

    @property
    def launch_count(self):
        return self._lib.swo_launch_count(self._h)

    def set_option(self, key, value):
        self._ck(self._lib.swo_set_option(self._h, key.encode(), int(value)))

    def tcid(self, contig):
        b = contig.encode() if isinstance(contig, str) else bytes(contig)
        return self._tcid.get(b, -1)

    def source_contig_size(self, tcid):
        return self._lib.swo_tcontig_size(self._h, tcid)

    def blocks(self, contig):
        """Blocks of one source contig in (tStart,tEnd,qStart) order as 5 numpy arrays."""
        t = contig if isinstance(contig, int) else self.tcid(contig)
        n = self._lib.swo_tcontig_nblocks(self._h, t)
        ts, te, qs, qc = (np.empty(n, np.int32) for _ in range(4))
        inv = np.empty(n, np.int8)
        self._ck(self._lib.swo_get_blocks(self._h, t, 0, n, _addr(ts), _addr(te), _addr(qs), _addr(qc), _addr(inv)))
        return ts, te, qs, qc, inv

    # -- batched operators ----------------------------------------------------------
    def lift_intervals(self, tcid, start, end, strand=None, thick_start=None, thick_end=None):
        """Batched ChainFile.lift + the numeric part of liftBED's record body
        (chain.d:664-713, bed.d:129-196).  Returns a dict of numpy arrays."""
        tcid = np.ascontiguousarray(tcid, np.int32)
        start = np.ascontiguousarray(start, np.int32)
        end = np.ascontiguousarray(end, np.int32)
        n = len(tcid)
        if not (len(start) == len(end) == n):
            raise ValueError("tcid/start/end lengths differ")
        strand = None if strand is None else np.ascontiguousarray(strand, np.uint8)
        ts = None if thick_start is None else np.ascontiguousarray(thick_start, np.int32)
        te = None if thick_end is None else np.ascontiguousarray(thick_end, np.int32)
        out = L.SwoLifted()
        self._ck(self._lib.swo_lift_intervals(self._h, n, _addr(tcid), _addr(start), _addr(end), _addr(strand),
                                              _addr(ts), _addr(te), C.byref(out)))
        nr = out.n_rows
        off = np.ctypeslib.as_array(out.row_offset, shape=(n + 1,)).copy()
        if nr:
            rows = np.ctypeslib.as_array(C.cast(out.rows, C.POINTER(C.c_int32)), shape=(nr, 4)).copy()
        else:
            rows = np.zeros((0, 4), np.int32)
        res = dict(n_rows=nr, n_unmatched=out.n_unmatched, row_offset=off, qcid=rows[:, 0] & 0xFFFF,
                   start=rows[:, 1], end=rows[:, 2], src=rows[:, 3].view(np.uint32))
        if strand is not None:
            res["strand"] = ((rows[:, 0] >> 16) & 0xFF).astype(np.uint8)
        if ts is not None:
            th = (np.ctypeslib.as_array(C.cast(out.thick, C.POINTER(C.c_int32)), shape=(nr, 2)).copy()
                  if nr else np.zeros((0, 2), np.int32))
            res["thick_start"], res["thick_end"] = th[:, 0], th[:, 1]
        return res

    def lift_points(self, tcid, pos):
        """Batched liftDirectly / liftCoordOnly (chain.d:604-654)."""
        tcid = np.ascontiguousarray(tcid, np.int32)
        pos = np.ascontiguousarray(pos, np.int32)
        n = len(tcid)
        oq, op, hit = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8)
        self._ck(self._lib.swo_lift_points(self._h, n, _addr(tcid), _addr(pos), _addr(oq), _addr(op), _addr(hit)))
        return oq, op, hit

    def lift_bed_text(self, data):
        """liftBED on an in-memory BED file (bed.d:72-204): returns (lifted, unmatched, stats)."""
        buf = bytes(data)
        out = L.SwoTextOut()
        self._ck(self._lib.swo_lift_bed_text(self._h, buf, len(buf), C.byref(out)))
        lifted = _bytes_at(out.lifted, out.lifted_len)
        unm = _bytes_at(out.unmatched, out.unmatched_len)
        return lifted, unm, dict(n_records=out.n_records, n_rows=out.n_rows, n_unmatched=out.n_unmatched)

    def lift_bed_text_ptr(self, ptr, nbytes):
        """Same, input given as a raw host address (e.g. pinned memory); returns the
        swo_text_out struct without copying the outputs (valid until the next call)."""
        out = L.SwoTextOut()
        self._ck(self._lib.swo_lift_bed_text(self._h, ptr, nbytes, C.byref(out)))
        return out

    def text_segments(self, which):
        """The last lift_bed_text result as per-shard (address, length) pieces in output order
        (which: 0 = lifted, 1 = unmatched) — see swo_text_segment."""
        out = []
        for i in range(self._lib.swo_text_num_segments(self._h)):
            p, n = C.c_void_p(), C.c_size_t()
            self._ck(self._lib.swo_text_segment(self._h, which, i, C.byref(p), C.byref(n)))
            out.append((p.value or 0, n.value))
        return out

    def load_genome_2bit(self, packed, contig_off, contig_len):
        packed = np.ascontiguousarray(packed, np.uint8)
        contig_off = np.ascontiguousarray(contig_off, np.int64)
        contig_len = np.ascontiguousarray(contig_len, np.int64)
        self._ck(self._lib.swo_load_genome_2bit(self._h, _addr(packed), packed.nbytes, _addr(contig_off),
                                                _addr(contig_len), len(contig_off)))

    def load_genome_fasta(self, path):
        """IndexedFastaFile(genomefile) (vcf.d:38): destination genome from an uncompressed FASTA."""
        self._ck(self._lib.swo_load_genome_fasta(self._h, path.encode()))

    def genome_seq_len(self, name):
        """fa.seqLen(name), -1 if the FASTA has no such sequence (fa.hasSeq, vcf.d:79-83)."""
        b = name.encode() if isinstance(name, str) else bytes(name)
        return self._lib.swo_genome_seq_len(self._h, b)

    def lift_vcf_text(self, data, chainfile="", genomefile="", filedate=None):
        """liftVCF on an in-memory text VCF (vcf.d:29-176): returns (lifted, unmatched, stats)."""
        buf = bytes(data)
        out = L.SwoTextOut()
        fd = None if filedate is None else filedate.encode()
        self._ck(self._lib.swo_lift_vcf_text(self._h, buf, len(buf), chainfile.encode(), genomefile.encode(), fd,
                                             C.byref(out)))
        lifted = _bytes_at(out.lifted, out.lifted_len)
        unm = _bytes_at(out.unmatched, out.unmatched_len)
        return lifted, unm, dict(n_records=out.n_records, n_matched=out.n_rows, n_unmatched=out.n_unmatched)

    def lift_maf_text(self, data, genomebuild):
        """liftMAF on an in-memory MAF (maf.d:119-319): returns (lifted, unmatched, stats)."""
        buf = bytes(data)
        out, st = L.SwoTextOut(), L.SwoMafStats()
        self._ck(self._lib.swo_lift_maf_text(self._h, buf, len(buf), genomebuild.encode(), C.byref(out), C.byref(st)))
        lifted = _bytes_at(out.lifted, out.lifted_len)
        unm = _bytes_at(out.unmatched, out.unmatched_len)
        return lifted, unm, {k: getattr(st, k) for k, _ in st._fields_}

    def lift_maf_records(self, tcid, start, end, reflen, ref_off, ref_bytes, out_off, out_cap):
        """Batched liftMAF record body (maf.d:189-291): zero-based half-open [start,end) in,
        one-based closed out; reflen -1 = "-" allele (no REF check)."""
        tcid = np.ascontiguousarray(tcid, np.int32)
        start = np.ascontiguousarray(start, np.int32)
        end = np.ascontiguousarray(end, np.int32)
        reflen = np.ascontiguousarray(reflen, np.int32)
        ref_off = np.ascontiguousarray(ref_off, np.uint64)
        out_off = np.ascontiguousarray(out_off, np.uint64)
        ref_bytes = np.ascontiguousarray(ref_bytes, np.uint8)
        n = len(tcid)
        nh, fl = np.empty(n, np.uint8), np.empty(n, np.uint8)
        oq, os_, oe, orl = (np.empty(n, np.int32) for _ in range(4))
        oref = np.zeros(max(int(out_cap), 1), np.uint8)
        self._ck(self._lib.swo_lift_maf_records(self._h, n, _addr(tcid), _addr(start), _addr(end), _addr(reflen),
                                                _addr(ref_off), _addr(ref_bytes) if len(ref_bytes) else None,
                                                len(ref_bytes), _addr(out_off), int(out_cap), _addr(nh), _addr(oq),
                                                _addr(os_), _addr(oe), _addr(fl), _addr(orl),
                                                _addr(oref) if out_cap else None))
        return dict(nhits=nh, qcid=oq, start=os_, end=oe, flags=fl, reflen=orl, ref=oref[:int(out_cap)])

    def lift_vcf_points(self, tcid, pos, reflen, ref_off, ref_bytes):
        """Batched liftVCF record body (vcf.d:115-169)."""
        tcid = np.ascontiguousarray(tcid, np.int32)
        pos = np.ascontiguousarray(pos, np.int32)
        reflen = np.ascontiguousarray(reflen, np.int32)
        ref_off = np.ascontiguousarray(ref_off, np.uint64)
        ref_bytes = np.ascontiguousarray(ref_bytes, np.uint8)
        n = len(tcid)
        oq, op, orl = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)
        fl = np.empty(n, np.uint8)
        oref = np.zeros(max(len(ref_bytes), 1), np.uint8)
        self._ck(self._lib.swo_lift_vcf_points(self._h, n, _addr(tcid), _addr(pos), _addr(reflen), _addr(ref_off),
                                               _addr(ref_bytes), len(ref_bytes), _addr(oq), _addr(op), _addr(fl),
                                               _addr(orl), _addr(oref)))
        return dict(qcid=oq, pos=op, flags=fl, reflen=orl, ref=oref[:len(ref_bytes)])

    # -- the reference's one-record interface ---------------------------------------
    def lift(self, contig, start, end, strand=None):
        """ChainFile.lift(contig,start,end) (chain.d:664): the pieces of [start,end) in
        destination coordinates, sorted as liftBED emits them (bed.d:156-157)."""
        s = None if strand is None else [strand[0] if isinstance(strand, (bytes, bytearray)) else ord(strand[0])]
        r = self.lift_intervals([self.tcid(contig)], [start], [end], s)
        return [LiftedInterval(self.contigNames[r["qcid"][k]], int(r["start"][k]), int(r["end"][k]),
                               bytes([r["strand"][k]]) if s is not None else b"") for k in range(r["n_rows"])]

    def liftCoordOnly(self, contig, coord):
        """chain.d:637-654: (nresults, coord')."""
        _, op, hit = self.lift_points([self.tcid(contig)], [coord])
        return int(hit[0]), int(op[0])

    def liftDirectly(self, contig, coord):
        """chain.d:604-627: (nresults, contig', coord')."""
        oq, op, hit = self.lift_points([self.tcid(contig)], [coord])
        name = contig.encode() if isinstance(contig, str) else contig
        return int(hit[0]), (self.contigNames[oq[0]] if hit[0] else name), int(op[0])
