"""CPU-side checks of the product library (no GPU needed, no compute calls):
  * libswiftover_cuda.so loads and exports every symbol include/*.h declares;
  * the host chain parser (swiftover_b200/csrc/chain_table.cpp, independent of the
    oracle) builds the same flattened block table as the oracle on every fixture and on
    seeded random chains, and rejects what the oracle rejects;
  * without a device the product fails loudly (no CPU fallback).
"""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

import synth
from conftest import ROOT, resource_bytes


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from swiftover_b200 import _lib
    return _lib


def declared_symbols():
    syms = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        syms |= set(re.findall(r"\b(swo_[a-z0-9_]+)\s*\(", txt))
    return syms


def test_library_exports_every_declared_symbol(built):
    L = C.CDLL(built.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in sorted(syms):
        assert hasattr(L, s), f"{s} declared in include/*.h but not exported"
    # and the Python binding knows every one of them
    assert syms <= set(built.lib()._signatures), syms - set(built.lib()._signatures)
    assert b"sm_100a" in built.lib().swo_version()


def test_no_cpu_fallback(built):
    import torch
    from swiftover_b200 import ChainFile, SwiftoverError
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    with pytest.raises(SwiftoverError) as e:
        ChainFile(text=resource_bytes("minimal.chain"))
    assert e.value.code == built.SWO_E_CUDA
    cf = ChainFile(text=resource_bytes("minimal.chain"), inspect_only=True)
    with pytest.raises(SwiftoverError) as e:
        cf.lift_intervals([0], [1], [2])
    assert e.value.code == built.SWO_E_CUDA
    with pytest.raises(SwiftoverError):
        cf.lift_bed_text(b"chr1\t1\t2\n")


def same_table(cf, oc):
    assert [n.decode() for n in cf.sourceContigs] == oc.tcontigs
    assert [n.decode() for n in cf.contigNames] == oc.qcontigs
    assert {k.decode(): v for k, v in cf.qContigSizes.items()} == oc.qsizes
    assert cf.num_blocks == oc.nblocks()
    for i, name in enumerate(oc.tcontigs):
        ts, te, qs, qc, inv = cf.blocks(i)
        ob = oc.blocks(i)
        assert len(ts) == len(ob)
        got = list(zip(ts.tolist(), te.tolist(), qs.tolist(), qc.tolist(), inv.tolist()))
        assert got == ob, name


@pytest.mark.parametrize("name", ["hg19ToHg38.over.chain", "chr1_first.chain", "minimal.chain"])
def test_chain_parser_matches_oracle_on_fixtures(built, oracle, name):
    from swiftover_b200 import ChainFile
    text = resource_bytes(name)
    cf = ChainFile(text=text, inspect_only=True)
    same_table(cf, oracle.chain_parse(text))
    assert cf.is_disjoint  # UCSC over.chain: no target overlaps
    if name.startswith("hg19"):
        assert cf.num_blocks == 53_950 and len(cf.sourceContigs) == 93
        assert cf.source_contig_size(cf.tcid("chr1")) == 249_250_621


@pytest.mark.parametrize("seed", range(8))
def test_chain_parser_matches_oracle_on_random_chains(built, oracle, seed):
    from swiftover_b200 import ChainFile
    rng = np.random.default_rng(50 + seed)
    text, _ = synth.random_chain(rng, n_chains=int(rng.integers(1, 15)), overlap=bool(seed % 2),
                                 sep=b" " if seed % 3 == 0 else b"\t")
    if seed == 5:  # duplicate the first chain: opCmp-equal blocks are one tree node
        first = text.split(b"\n\n")[0]
        text = text + first + b"\n"
    cf = ChainFile(text=text, inspect_only=True)
    oc = oracle.chain_parse(text)
    same_table(cf, oc)
    disjoint = True
    for i in range(len(oc.tcontigs)):
        pm = None
        for b in oc.blocks(i):
            if pm is not None and b[0] < pm:
                disjoint = False
            pm = b[1] if pm is None else max(pm, b[1])
    assert cf.is_disjoint == disjoint
    if not seed % 2 and seed != 5:
        assert disjoint


def test_chain_parser_file_and_errors(built, oracle, resources, tmp_path):
    from swiftover_b200 import ChainFile, SwiftoverError
    from oracle_ffi import OracleError
    cf = ChainFile(os.path.join(resources, "minimal.chain"), inspect_only=True)
    assert cf.num_blocks == 9
    with pytest.raises(SwiftoverError) as e:
        ChainFile(str(tmp_path / "missing.chain"), inspect_only=True)
    assert e.value.code == built.SWO_E_IO  # FileException, chain.d:513-514
    bad = [b"", b"chain 1 chr1 10 + 0 5\n5\n", b"chain 1 a 100 + 0 50 b 100 + 0 50 1\n10 5\n",
           b"chain 1 a 100 + 0 50 b 100 + 0 50 1\n10 x 5\n", b"# c\n# d\nchain 1 a 100 + 0 50 b 100 + 0 50 1\n50\n"]
    for text in bad:
        with pytest.raises(SwiftoverError) as e:
            ChainFile(text=text, inspect_only=True)
        assert e.value.code == built.SWO_E_PARSE
        with pytest.raises(OracleError):
            oracle.chain_parse(text)
    # quirks both sides reproduce: one leading junk line is ignored (chain.d:528-530) and a
    # chain that ends at line index 0 is dropped
    for text in (b"junk\nchain 1 a 100 + 0 50 b 100 + 0 50 1\n50\n",
                 b"chain 9 z 100 + 0 50 y 100 + 0 50 9\nchain 1 a 100 + 0 50 b 100 + 0 50 1\n50\n",
                 b"chain 1 a 100 - 10 60 b 100 - 0 50 1\n20 5 5\n25\n\n"):
        same_table(ChainFile(text=text, inspect_only=True), oracle.chain_parse(text))


def test_python_wrapper_copies_outputs_by_size_t():
    """Outputs larger than 2 GiB must survive the ctypes layer (ctypes.string_at takes a C int)."""
    import ctypes as C
    from swiftover_b200.chain import _bytes_at
    buf = C.create_string_buffer(b"lifted\tbytes\n", 13)
    assert _bytes_at(C.addressof(buf), 13) == b"lifted\tbytes\n" and _bytes_at(0, 0) == b""
    assert _bytes_at(C.addressof(buf), C.c_size_t(6).value) == b"lifted"
