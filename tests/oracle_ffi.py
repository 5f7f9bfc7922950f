"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference algorithm (oracle/swo_oracle.h).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product (swiftover_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")


def build_oracle():
    src = [os.path.join(ORACLE_DIR, f) for f in ("swo_oracle.c", "swo_oracle.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "liboracle.so", "orc_cli"])



def _bytes_at(ptr, n):
    """ctypes.string_at takes a C int size: fails past 2 GiB."""
    if not n:
        return b""
    addr = ptr if isinstance(ptr, int) else C.cast(ptr, C.c_void_p).value
    return bytes((C.c_char * n).from_address(addr))

class OrcLink(C.Structure):
    _fields_ = [("tStart", C.c_int64), ("tEnd", C.c_int64), ("qStart", C.c_int64),
                ("qcid", C.c_int32), ("invert", C.c_int32)]

    def tup(self):
        return (self.tStart, self.tEnd, self.qStart, self.qcid, self.invert)


class OrcBuf(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t)]


class OrcBedStats(C.Structure):
    _fields_ = [("n_lines", C.c_uint64), ("n_records", C.c_uint64), ("n_rows", C.c_uint64),
                ("n_unmatched", C.c_uint64)]


class OrcRows(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("n_unmatched", C.c_uint64),
                ("row_offset", C.POINTER(C.c_uint32)), ("qcid", C.POINTER(C.c_int32)),
                ("start", C.POINTER(C.c_int32)), ("end", C.POINTER(C.c_int32)),
                ("src", C.POINTER(C.c_uint32)), ("strand", C.POINTER(C.c_uint8)),
                ("thick_start", C.POINTER(C.c_int32)), ("thick_end", C.POINTER(C.c_int32)),
                ("unmatched_idx", C.POINTER(C.c_uint32))]


def _np(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


class Oracle:
    def __init__(self):
        L = self.L = C.CDLL(LIB)
        L.orc_chain_parse.restype = C.c_void_p
        L.orc_chain_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        L.orc_chain_load.restype = C.c_void_p
        L.orc_chain_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.orc_chain_free.argtypes = [C.c_void_p]
        L.orc_chain_count_links.restype = C.c_int64
        L.orc_chain_count_links.argtypes = [C.c_char_p, C.c_size_t]
        for f in ("orc_num_tcontigs", "orc_num_qcontigs", "orc_num_qsizes"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("orc_tcontig_name", "orc_qcontig_name", "orc_qsize_name"):
            getattr(L, f).restype = C.c_char_p
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        L.orc_tcontig_nblocks.restype = C.c_int64
        L.orc_tcontig_nblocks.argtypes = [C.c_void_p, C.c_int]
        L.orc_total_blocks.restype = C.c_int64
        L.orc_total_blocks.argtypes = [C.c_void_p]
        L.orc_qsize_value.restype = C.c_int64
        L.orc_qsize_value.argtypes = [C.c_void_p, C.c_int]
        L.orc_tcontig_id.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.orc_tcontig_block.restype = OrcLink
        L.orc_tcontig_block.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        L.orc_lift.restype = C.c_int64
        L.orc_lift.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.POINTER(OrcLink))]
        L.orc_lift_coord_only.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.orc_lift_directly.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
        L.orc_lift_bed.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.POINTER(OrcBuf), C.POINTER(OrcBuf),
                                   C.POINTER(OrcBedStats), C.c_char_p, C.c_size_t]
        L.orc_lift_bed_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(OrcBuf),
                                      C.POINTER(OrcBuf), C.POINTER(OrcBedStats), C.c_char_p, C.c_size_t]
        L.orc_buf_free.argtypes = [C.POINTER(OrcBuf)]
        L.orc_lift_intervals.argtypes = [C.c_void_p, C.c_size_t] + [C.c_void_p] * 6 + [C.POINTER(OrcRows)]
        L.orc_rows_free.argtypes = [C.POINTER(OrcRows)]
        L.orc_lift_vcf_points.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_size_t] + [C.c_void_p] * 10
        L.orc_format_bed3.restype = C.c_size_t
        L.orc_format_bed3.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_link_qend.restype = C.c_int64
        L.orc_link_qend.argtypes = [C.POINTER(OrcLink)]
        L.orc_link_size.restype = C.c_int64
        L.orc_link_size.argtypes = [C.POINTER(OrcLink)]
        L.orc_link_cmp.restype = C.c_int64
        L.orc_link_cmp.argtypes = [C.POINTER(OrcLink), C.POINTER(OrcLink)]
        L.orc_link_cmp_scalar.argtypes = [C.POINTER(OrcLink), C.c_int64]
        L.orc_intersect.restype = OrcLink
        L.orc_intersect.argtypes = [OrcLink, C.c_int64, C.c_int64]
        L.orc_strand_flip.restype = C.c_ubyte
        L.orc_strand_flip.argtypes = [C.c_ubyte, C.c_int]

    # -- chain ---------------------------------------------------------
    def chain_parse(self, text: bytes):
        err = C.create_string_buffer(256)
        h = self.L.orc_chain_parse(text, len(text), err, 256)
        if not h:
            raise OracleError(-2, err.value.decode())
        return OracleChain(self, h)

    def chain_load(self, path: str):
        err = C.create_string_buffer(256)
        h = self.L.orc_chain_load(path.encode(), err, 256)
        if not h:
            raise OracleError(-1, err.value.decode())
        return OracleChain(self, h)

    def reverse_complement(self, allele: bytes) -> bytes:
        """reverse_complement (maf.d:36-54)"""
        out = C.create_string_buffer(len(allele) + 1)
        fn = self.L.orc_reverse_complement
        fn.restype = C.c_size_t
        fn.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        n = fn(allele, len(allele), out)
        return out.raw[:n]

    def count_links(self, text: bytes) -> int:
        return self.L.orc_chain_count_links(text, len(text))


class OracleChain:
    def __init__(self, o: Oracle, h):
        self.o, self.L, self.h = o, o.L, h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_chain_free(self.h)
            self.h = None

    @property
    def tcontigs(self):
        return [self.L.orc_tcontig_name(self.h, i).decode() for i in range(self.L.orc_num_tcontigs(self.h))]

    @property
    def qcontigs(self):
        return [self.L.orc_qcontig_name(self.h, i).decode() for i in range(self.L.orc_num_qcontigs(self.h))]

    @property
    def qsizes(self):
        return {self.L.orc_qsize_name(self.h, i).decode(): self.L.orc_qsize_value(self.h, i)
                for i in range(self.L.orc_num_qsizes(self.h))}

    def nblocks(self, i=None):
        return self.L.orc_total_blocks(self.h) if i is None else self.L.orc_tcontig_nblocks(self.h, i)

    def tcid(self, name) -> int:
        b = name.encode() if isinstance(name, str) else name
        return self.L.orc_tcontig_id(self.h, b, len(b))

    def blocks(self, i):
        return [self.L.orc_tcontig_block(self.h, i, j).tup() for j in range(self.nblocks(i))]

    def lift(self, contig, start, end):
        """ChainFile.lift (chain.d:664): list of (tStart,tEnd,qStart,qcid,invert), hit order."""
        p = C.POINTER(OrcLink)()
        tc = contig if isinstance(contig, int) else self.tcid(contig)
        n = self.L.orc_lift(self.h, tc, start, end, C.byref(p))
        out = [p[i].tup() for i in range(n)]
        C.CDLL(None).free(p)
        return out

    def lift_coord_only(self, contig, coord):
        tc = contig if isinstance(contig, int) else self.tcid(contig)
        c = C.c_int64(coord)
        r = self.L.orc_lift_coord_only(self.h, tc, C.byref(c))
        return r, c.value

    def lift_directly(self, contig, coord):
        tc = contig if isinstance(contig, int) else self.tcid(contig)
        c = C.c_int64(coord)
        q = C.c_int32(-1)
        r = self.L.orc_lift_directly(self.h, tc, C.byref(q), C.byref(c))
        return r, (self.qcontigs[q.value] if r else None), c.value

    def lift_bed(self, text: bytes, threads: int = 0):
        """liftBED (bed.d:72) text -> (lifted, unmatched, stats)."""
        lifted, unm, st = OrcBuf(), OrcBuf(), OrcBedStats()
        err = C.create_string_buffer(256)
        if threads and threads > 0:
            buf = C.create_string_buffer(text, len(text))
            rc = self.L.orc_lift_bed_mt(self.h, buf, len(text), threads, C.byref(lifted), C.byref(unm),
                                        C.byref(st), err, 256)
        else:
            rc = self.L.orc_lift_bed(self.h, text, len(text), C.byref(lifted), C.byref(unm), C.byref(st), err, 256)
        a = _bytes_at(lifted.data, lifted.len)
        b = _bytes_at(unm.data, unm.len)
        self.L.orc_buf_free(C.byref(lifted))
        self.L.orc_buf_free(C.byref(unm))
        if rc:
            raise OracleError(rc, err.value.decode())
        return a, b, dict(n_lines=st.n_lines, n_records=st.n_records, n_rows=st.n_rows, n_unmatched=st.n_unmatched)

    def lift_bed_ptr(self, ptr, n, threads):
        """Timing entry for bench.py: input already in memory at `ptr`; outputs discarded."""
        lifted, unm, st = OrcBuf(), OrcBuf(), OrcBedStats()
        err = C.create_string_buffer(256)
        rc = self.L.orc_lift_bed_mt(self.h, ptr, n, threads, C.byref(lifted), C.byref(unm), C.byref(st), err, 256)
        res = dict(lifted_len=lifted.len, unmatched_len=unm.len, n_records=st.n_records, n_rows=st.n_rows,
                   n_unmatched=st.n_unmatched)
        self.L.orc_buf_free(C.byref(lifted))
        self.L.orc_buf_free(C.byref(unm))
        if rc:
            raise OracleError(rc, err.value.decode())
        return res

    def format_bed3(self, tcid, start, end):
        """bench helper: BED3 text (numpy uint8 array) of the records."""
        tcid = np.ascontiguousarray(tcid, np.int32)
        start = np.ascontiguousarray(start, np.int32)
        end = np.ascontiguousarray(end, np.int32)
        out = np.empty(len(tcid) * 64 + 64, np.uint8)
        m = self.L.orc_format_bed3(self.h, len(tcid), tcid.ctypes.data, start.ctypes.data, end.ctypes.data,
                                   out.ctypes.data, out.nbytes)
        return out[:m]

    def tcontig_size_from(self, sizes_by_name):
        return [sizes_by_name[n] for n in self.tcontigs]

    def lift_intervals(self, tcid, start, end, strand=None, thick_start=None, thick_end=None):
        tcid = np.ascontiguousarray(tcid, np.int32)
        start = np.ascontiguousarray(start, np.int32)
        end = np.ascontiguousarray(end, np.int32)
        strand = None if strand is None else np.ascontiguousarray(strand, np.uint8)
        ts = None if thick_start is None else np.ascontiguousarray(thick_start, np.int32)
        te = None if thick_end is None else np.ascontiguousarray(thick_end, np.int32)
        n = len(tcid)
        r = OrcRows()
        vp = lambda a: a.ctypes.data if a is not None else None
        rc = self.L.orc_lift_intervals(self.h, n, vp(tcid), vp(start), vp(end), vp(strand), vp(ts), vp(te), C.byref(r))
        if rc:
            raise OracleError(rc, "lift_intervals")
        nr, nu = r.n_rows, r.n_unmatched
        out = dict(n_rows=nr, n_unmatched=nu, row_offset=_np(r.row_offset, n + 1, np.uint32),
                   qcid=_np(r.qcid, nr, np.int32), start=_np(r.start, nr, np.int32), end=_np(r.end, nr, np.int32),
                   src=_np(r.src, nr, np.uint32), unmatched_idx=_np(r.unmatched_idx, nu, np.uint32))
        if strand is not None:
            out["strand"] = _np(r.strand, nr, np.uint8)
        if ts is not None:
            out["thick_start"] = _np(r.thick_start, nr, np.int32)
            out["thick_end"] = _np(r.thick_end, nr, np.int32)
        self.L.orc_rows_free(C.byref(r))
        return out

    def lift_vcf_text(self, fasta: bytes, text: bytes, chainfile="", genomefile="", filedate="20240101"):
        """liftVCF on text (vcf.d:29-176): (lifted, unmatched, stats)."""
        class St(C.Structure):
            _fields_ = [("n_records", C.c_uint64), ("n_matched", C.c_uint64), ("n_unmatched", C.c_uint64),
                        ("n_refchg", C.c_uint64)]
        fn = self.L.orc_lift_vcf_text
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p,
                       C.POINTER(OrcBuf), C.POINTER(OrcBuf), C.POINTER(St), C.c_char_p, C.c_size_t]
        lifted, unm, st = OrcBuf(), OrcBuf(), St()
        err = C.create_string_buffer(256)
        rc = fn(self.h, fasta, len(fasta), text, len(text), chainfile.encode(), genomefile.encode(), filedate.encode(),
                C.byref(lifted), C.byref(unm), C.byref(st), err, 256)
        out = (_bytes_at(lifted.data, lifted.len),
               _bytes_at(unm.data, unm.len),
               dict(n_records=st.n_records, n_matched=st.n_matched, n_unmatched=st.n_unmatched, n_refchg=st.n_refchg))
        self.L.orc_buf_free(C.byref(lifted))
        self.L.orc_buf_free(C.byref(unm))
        if rc:
            raise OracleError(rc, err.value.decode())
        return out

    def lift_maf_text(self, fasta: bytes, text: bytes, genomebuild="B"):
        """liftMAF on text (maf.d:119-319): (lifted, unmatched, stats)."""
        keys = ("n_records", "n_matched", "n_unmatched", "n_multiple", "n_refchg", "n_neither")

        class St(C.Structure):
            _fields_ = [(k, C.c_uint64) for k in keys]
        fn = self.L.orc_lift_maf_text
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.POINTER(OrcBuf),
                       C.POINTER(OrcBuf), C.POINTER(St), C.c_char_p, C.c_size_t]
        lifted, unm, st = OrcBuf(), OrcBuf(), St()
        err = C.create_string_buffer(256)
        rc = fn(self.h, fasta, len(fasta), text, len(text), genomebuild.encode(), C.byref(lifted), C.byref(unm),
                C.byref(st), err, 256)
        out = (_bytes_at(lifted.data, lifted.len),
               _bytes_at(unm.data, unm.len), {k: getattr(st, k) for k in keys})
        self.L.orc_buf_free(C.byref(lifted))
        self.L.orc_buf_free(C.byref(unm))
        if rc:
            raise OracleError(rc, err.value.decode())
        return out

    def lift_vcf_points(self, packed, contig_off, contig_len, tcid, pos, reflen, ref_off, ref_bytes):
        packed = np.ascontiguousarray(packed, np.uint8)
        contig_off = np.ascontiguousarray(contig_off, np.int64)
        contig_len = np.ascontiguousarray(contig_len, np.int64)
        tcid = np.ascontiguousarray(tcid, np.int32)
        pos = np.ascontiguousarray(pos, np.int32)
        reflen = np.ascontiguousarray(reflen, np.int32)
        ref_off = np.ascontiguousarray(ref_off, np.uint64)
        ref_bytes = np.ascontiguousarray(ref_bytes, np.uint8)
        n = len(tcid)
        oq, op = np.empty(n, np.int32), np.empty(n, np.int32)
        fl, orl = np.empty(n, np.uint8), np.empty(n, np.int32)
        oref = np.zeros(max(len(ref_bytes), 1), np.uint8)
        rc = self.L.orc_lift_vcf_points(self.h, packed.ctypes.data, contig_off.ctypes.data, contig_len.ctypes.data,
                                        len(contig_off), n, tcid.ctypes.data, pos.ctypes.data, reflen.ctypes.data,
                                        ref_off.ctypes.data, ref_bytes.ctypes.data, oq.ctypes.data, op.ctypes.data,
                                        fl.ctypes.data, orl.ctypes.data, oref.ctypes.data)
        if rc:
            raise OracleError(rc, "lift_vcf_points")
        return dict(qcid=oq, pos=op, flags=fl, reflen=orl, ref=oref[:len(ref_bytes)])
