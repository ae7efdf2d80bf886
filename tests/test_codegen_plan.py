"""Host-side checks of the chain code generator (no GPU): the ICPC recipe is planned on the
"meta" device, and the generated kernel is inspected -- node order, shared-memory budget,
the two instruction streams and their events -- and compiled for sm_100a."""

import os
import re
import shutil

import numpy as np
import pytest
import yaml

from dspeed_b200 import codegen

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ICPC = os.path.join(REPO, "dspeed_b200", "configs", "hpge_icpc.yaml")


@pytest.fixture(scope="module")
def spec():
    import numpy as np

    from dspeed_b200 import tables
    from dspeed_b200.processing_chain import build_processing_chain

    n = 4
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=np.zeros((n, 8192), np.uint16))
    tb = tables.Table({"waveform": wf, "baseline": tables.Array(np.zeros(n, np.uint16))}, size=n)
    chain, _, _ = build_processing_chain(yaml.safe_load(open(ICPC)), tb, block_width=16, device="meta")
    build = codegen.SpecChain._build
    codegen.SpecChain._build = lambda self: None     # plan only
    try:
        return codegen.SpecChain(chain)
    finally:
        codegen.SpecChain._build = build


def test_icpc_program(spec):
    kinds = [nd["kind"] for nd in spec.order]
    # the raw row is loaded once; filters of one input are evaluated together; cusp + zac share
    # one evaluation; the trapezoid that is only picked off is never materialised
    assert kinds.count("load") == 1 and kinds.count("fir_group") == 1 and kinds.count("conv_seg_group") == 1
    # 11 threshold searches: tp_95 ... tp_01 (each starting at the previous result) are one combined walk
    assert kinds.count("fir_lazy") == 1 and kinds.count("tpt") == 4 and kinds.count("tpt_chain") == 1
    chain = next(nd for nd in spec.order if nd["kind"] == "tpt_chain")
    assert len(chain["members"]) == 7 and chain["start"] == next(nd["out"] for nd in spec.order if nd["kind"] == "tpt" and
                                                               nd["out"] == chain["start"])
    assert [c[0] for c in spec.conv_lowering] == ["runs", "seg", "seg"] and spec.cse_skipped == 1
    # reductions follow their producer (they read the chunk from registers)
    assert kinds[:4] == ["load", "min_max", "bl_sub", "lsf"]
    assert spec.smem_bytes <= 227 * 1024 and spec.total_slots % 2 == 0
    assert len(spec.out_scalars) == 34


def _block_stream(src, idle=False):
    """block-stream text; the idle branches of the short-waveform regions (warps without a chunk of those waveforms
    only mirror the synchronisation) are cut out (or returned alone)"""
    blk = src[src.index("block stream (warps 0-15)"):src.index("scalar stream (warp 16)")]
    pat = re.compile(r"\} else \{   // warps without a chunk.*?\n\s*\}\n", re.S)
    if idle:
        return "\n".join(pat.findall(blk))
    return pat.sub("}\n", blk)


def _scalar_streams(src):
    """texts of the scalar warps' streams (warp 16, and warp 17 when the scalar work is split)"""
    a = src.index("scalar stream (warp 16)")
    b = src.find("second scalar stream (warp 17)")
    end = src.index("#undef EVB")
    return [src[a:b], src[b:end]] if b >= 0 else [src[a:end]]


def test_streams_and_events(spec):
    src = spec.source()
    blk = _block_stream(src)
    streams = _scalar_streams(src)
    sca = streams[0]
    # threshold searches, pick-offs and output stores live in the scalar streams only
    assert "tpt_w(" in sca and "tpt_w(" not in blk
    assert all("A.p[" in st and "st_chunk_n" not in st for st in streams)
    # every block -> scalar event has exactly one arrive and one wait per scalar warp, in the same order
    arr = re.findall(r"EV_ARRIVE\(EVB\((\d+)\)\)", blk)
    for st in streams:
        assert re.findall(r"EV_WAIT\(EVB\((\d+)\)\)", st) == arr
    assert 1 <= len(arr) <= 5
    # the scalar published to the block stream (tp_0_est -> windower) crosses once, from the first scalar warp
    assert blk.count("EV_WAIT(0);") == 1 and sca.count("EV_ARRIVE(0);") == 1
    assert all("EV_ARRIVE(0);" not in st for st in streams[1:])
    # cross-row pipelining: "done" event per row and scalar warp, consumed before the next row's first colliding write
    assert all(st.count("EV_ARRIVE(15);") == 1 for st in streams) and "if (it > 0) EV_WAIT(15);" in blk
    assert len({st.count("EV_ARRIVE(14);") for st in streams}) == 1
    # block-only barriers never involve the scalar warps
    assert "__syncthreads()" not in blk


def test_scalar_work_is_split_over_two_warps(spec):
    """the ICPC chain's 11 dependent threshold searches are the longer pipeline: the path to tp_0_est stays on the first
    scalar warp, whole dependency components of the rest go to the second one, which receives tp_0_est exactly once"""
    assert spec.n_swarps == 2
    src = spec.source()
    s1, s2 = _scalar_streams(src)
    assert s1.count("tpt_w(") + s2.count("tpt_w(") == 4 and (s1 + s2).count("tpt_chain_bwd<7>(") == 1
    assert "tpt_w(" in s1 and ("tpt_w(" in s2 or "tpt_chain_bwd" in s2)
    assert s1.count("EV_ARRIVE(EVX);") == 1 and s2.count("EV_WAIT(EVX);") == 1 and "EV_ARRIVE(EVX)" not in s2
    # the hand-over follows the search that produces tp_0_est and precedes every use in the second warp
    assert s1.index("EV_ARRIVE(0);") < s1.index("EV_ARRIVE(EVX);")
    assert s2.index("EV_WAIT(EVX);") < min(s2.find(x) for x in ("tpt_w(", "tpt_chain_bwd") if x in s2)
    # every output column is stored exactly once
    stores = re.findall(r"\(\(float\*\)A\.p\[(\d+)\]\)\[row\] = ", s1 + s2)
    assert len(stores) == len(set(stores)) == len(spec.out_scalars)
    # launch: 16 block warps + 2 scalar warps, events counted accordingly
    assert "__launch_bounds__(576, 1)" in src and "<<<grid, 576," in src
    assert "? 64 : 576)" in src


def test_idle_warps_of_a_region_mirror_its_synchronisation(spec):
    """warps that own no chunk of a region's short waveforms execute the same barriers / events in the same order"""
    src = spec.source()
    blk = src[src.index("block stream (warps 0-15)"):src.index("scalar stream (warp 16)")]
    regions = re.findall(r"// ---- region:.*?\n(.*?)\} else \{   // warps without a chunk[^\n]*\n(.*?)\n\s*\}\n", blk, re.S)
    assert regions, "the ICPC chain has a short-waveform region (windowed current)"
    sync = re.compile(r"BSYNC\(\)|EV_ARRIVE\([^)]*\)\)?|EV_WAIT\([^)]*\)\)?|par \^= 1")
    for body, idle in regions:
        assert sync.findall(body) == sync.findall(idle)
        assert "ld_shift" not in idle and "st_chunk_n" not in idle


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_generated_kernel_compiles_for_sm100a(spec):
    so, cu = codegen.build_source(spec.source())
    assert os.path.exists(so) and os.path.getsize(so) > 10000
    assert "sm_100a" in " ".join(codegen._lib.NVCC_FLAGS)


def test_done_event_wait_precedes_the_last_block_to_scalar_event(spec):
    """the block stream's wait for "scalar warp done with the previous row" must come before the row's last
    block -> scalar event (else a fast scalar warp can arrive twice in one barrier phase: deadlock)"""
    src = spec.source()
    blk = _block_stream(src)
    assert blk.index("if (it > 0) EV_WAIT(15);") < blk.rindex("EV_ARRIVE(EVB(")


def test_double_pole_zero_is_specialised():
    """double_pole_zero with constant time constants has an emitter (geometric scan + plain scan, two barrier rounds)"""
    import numpy as np

    from dspeed_b200 import tables
    from dspeed_b200.processing_chain import build_processing_chain

    cfg = {"outputs": ["dpz_max"], "processors": {
        "wf_blsub": "dspeed.processors.bl_subtract(waveform, baseline, wf_blsub(unit='ADC'))",
        "wf_dpz": {"function": "dspeed.processors.double_pole_zero(wf_blsub, 27460.5, 1200.25, 0.025, wf_dpz)", "unit": "ADC"},
        "dpz_max": {"function": "numpy.amax(wf_dpz, 1, dpz_max)", "unit": "ADC"}}}
    n = 4
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=np.zeros((n, 8192), np.uint16))
    tb = tables.Table({"waveform": wf, "baseline": tables.Array(np.zeros(n, np.uint16))}, size=n)
    chain, _, _ = build_processing_chain(cfg, tb, block_width=16, device="meta")
    build = codegen.SpecChain._build
    codegen.SpecChain._build = lambda self: None
    try:
        sc = codegen.SpecChain(chain)
    finally:
        codegen.SpecChain._build = build
    assert [nd["kind"] for nd in sc.order][:3] == ["load", "bl_sub", "dpz"]
    src = sc.source()
    assert "dpz_local(" in src and "put_scan_geo(" in src and "get_excl_geo(" in src
    # the row's only block -> scalar event comes after the wait for the previous row's "done" event
    blk = src[src.index("block stream (warps 0-15)"):src.index("scalar stream (warp 16)")]
    assert blk.index("if (it > 0) EV_WAIT(15);") < blk.rindex("EV_ARRIVE(EVB(")


def test_database_constants_do_not_change_the_kernel():
    """chains that differ only in a per-channel / per-run database value (the pole-zero time constant) share ONE compiled
    kernel: the value travels through the launch arguments (Args.c[]), not through the source text"""
    import yaml

    from dspeed_b200 import tables
    from dspeed_b200.processing_chain import build_processing_chain

    n = 4
    srcs, consts = [], []
    for tau in ("439.368*us", "431.2*us"):
        cfg = yaml.safe_load(open(os.path.join(os.path.dirname(codegen.__file__), "configs", "hpge_icpc.yaml")))
        wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=np.zeros((n, 8192), np.uint16))
        tb = tables.Table({"waveform": wf, "baseline": tables.Array(np.zeros(n, np.uint16))}, size=n)
        # (the whole chain: pole_zero's 1 - exp(-1 / tau) and the exp(-1 / tau) of the cusp / zac kernel model)
        chain, _, _ = build_processing_chain(cfg, tb, db_dict={"pz": {"tau": tau}}, block_width=16, device="meta")
        build = codegen.SpecChain._build
        codegen.SpecChain._build = lambda self: None
        try:
            sc = codegen.SpecChain(chain)
        finally:
            codegen.SpecChain._build = build
        srcs.append(sc.source())
        consts.append(list(sc.rt_consts))
    assert srcs[0] == srcs[1] and "A.c[0]" in srcs[0]
    assert consts[0] != consts[1] and len(consts[0]) == 2
    for cs, tau in zip(consts, (439368.0 / 16, 431200.0 / 16)):
        tau32 = np.float64(np.float32(tau))
        assert sorted(abs(c - v) < 1e-15 for c, v in zip(sorted(cs), sorted([1 - np.exp(-1 / tau32), np.exp(-1 / tau32)]))) == [True, True]


def test_small_chains_get_two_resident_ctas_per_sm(spec):
    """A chain that needs few shared-memory slots is planned a second time with the pool cut to twice its logical slots,
    so that two CTAs fit one SM (launch bounds (threads, 2), grid = 2 x SMs); the ICPC chain (6 slots of 32.9 KB) keeps
    one CTA per SM.  The second planning pass starts from a clean state (the raw-chunk prefetch is still emitted)."""
    from dspeed_b200 import tables
    from dspeed_b200.processing_chain import build_processing_chain
    from scripts.bench_configs import C1_CFG

    n = 4
    wf = tables.WaveformTable(size=n, t0=0, t0_units="ns", dt=16, dt_units="ns", values=np.zeros((n, 8192), np.uint16))
    chain, _, _ = build_processing_chain(C1_CFG, tables.Table({"waveform": wf}, size=n), block_width=16, device="meta")
    sc = codegen.SpecChain(chain)
    src = sc.source()
    assert sc.occ == 2 and sc.smem_bytes <= (codegen.MAX_SMEM - 2048) // 2
    assert f"__launch_bounds__({512 + 32 * sc.n_swarps}, 2)" in src and "num_sms * 2" in src
    assert src.count("pf0 = ldg_nc(g_)") == 2          # prologue + in-loop prefetch of the next row
    assert codegen.spill_bytes(sc.lib_path) == 0
    assert spec.smem_bytes > (codegen.MAX_SMEM - 2048) // 2 and "__launch_bounds__(576, 1)" in spec.source()
