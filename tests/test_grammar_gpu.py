"""The reference's expression-grammar tests (tests/test_processing_chain.py:289-620 of legend-exp/dspeed) executed on
the device through build_dsp: unit conversion of processor arguments, coordinates on several grids (windowed and
strided views), round / floor / ceil / trunc (including round-half-to-even), `where` / ternary with the unit
reconciliation rules, isnan / isfinite, astype.  Expected values are the reference tests' own assertions; the raw table
is synthetic (the reference reads a LEGEND test file): 8000-sample uint16 waveforms on a 16 ns grid with a pulse, plus
an `eventnumber` column."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def raw_tbl():
    from dspeed_b200 import tables

    n, L = 4, 8000
    rng = np.random.default_rng(11)
    t = np.arange(L)
    vals = np.empty((n, L), np.uint16)
    for r in range(n):
        x = np.clip(t - (3250 + 40 * r), 0, None)
        w = 12000 + 60 * r + (4000 + 900 * r) * (1 - np.exp(-x / 12.0)) * np.exp(-x / 27000.0) + rng.normal(0, 4, L)
        vals[r] = np.clip(np.rint(w), 0, 65535)
    wf = tables.WaveformTable(size=n, t0=tables.Array(np.array([0.0, 32.0, 64.0, 96.0]), attrs={"units": "ns"}),
                              dt=16, dt_units="ns", values=vals)
    return tables.Table({"waveform": wf, "eventnumber": tables.Array(np.arange(n, dtype=np.int32))}, size=n)


def build(raw_tbl, cfg, **kw):
    from dspeed_b200.build_dsp import build_dsp

    return build_dsp(raw_tbl, dsp_config=cfg, **kw)


def col(out, name):
    a = out[name].nda
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def wvals(out, name):
    a = out[name].values.nda
    return a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)


def test_unit_conversion_of_arguments(raw_tbl):
    """:289-320 -- 100 samples = 1600 ns = 1.6 us = 6.25 GHz x 16 ns on a 16 ns grid"""
    P = {"function": "fixed_time_pickoff", "module": "dspeed.processors"}
    cfg = {"outputs": ["a_unitless", "a_ns", "a_us", "a_ghz"], "processors": {
        "a_unitless": {**P, "args": ["waveform", 100, "'n'", "a_unitless"]},
        "a_ns": {**P, "args": ["waveform", "1600*ns", "'n'", "a_ns"]},
        "a_us": {**P, "args": ["waveform", "1.6*us", "'n'", "a_us"]},
        "a_ghz": {**P, "args": ["waveform", "6.25*GHz", "'n'", "a_ghz"]}}}
    out = build(raw_tbl, cfg, n_entries=1)
    v = raw_tbl["waveform"].values.nda[0, 100]
    assert col(out, "a_unitless")[0] == v == col(out, "a_ns")[0] == col(out, "a_us")[0] == col(out, "a_ghz")[0]


def test_coordinate_grids(raw_tbl):
    """:326-390 -- a time picked off a windowed view and a down-sampled view of the waveform"""
    F, T = ({"function": "fixed_time_pickoff", "module": "dspeed.processors"},
            {"function": "time_point_thresh", "module": "dspeed.processors"})
    cfg = {"outputs": ["a_window", "a_downsample", "tp", "tp_window", "tp_downsample"], "processors": {
        "a_window": {**F, "args": ["waveform[2625:4025]", "51.2*us + waveform.offset", "'i'", "a_window"], "unit": ["ADC"]},
        "a_downsample": {**F, "args": ["waveform[0:8000:8]", "51.2*us + waveform.offset", "'i'", "a_downsample"], "unit": ["ADC"]},
        "tp": {**T, "args": ["waveform", "a_window", "52.48*us+waveform.offset", 0, "tp"], "unit": "ns"},
        "tp_window": {**T, "args": ["waveform[2625:4025]", "a_window", "52.48*us+waveform.offset", 0, "tp_window"], "unit": "ns"},
        "tp_downsample": {**T, "args": ["waveform[0:8000:8]", "a_window", "52.48*us+waveform.offset", 0, "tp_downsample"], "unit": "ns"}}}
    out = build(raw_tbl, cfg, n_entries=2)
    w = raw_tbl["waveform"].values.nda
    for r in range(2):
        assert col(out, "a_window")[r] == col(out, "a_downsample")[r] == w[r, 3200]      # 51.2 us = sample 3200
        assert col(out, "tp_window")[r] == col(out, "tp")[r]
        assert -128 < col(out, "tp_downsample")[r] - col(out, "tp")[r] < 128
        # backward from sample 3280 for the first crossing of the value at sample 3200, in ns incl. the event's t0
        i = 3280
        thr = np.float32(w[r, 3200])
        x = w[r].astype(np.float32)
        while i >= 1 and not ((x[i - 1] < thr <= x[i]) or (x[i - 1] > thr >= x[i])):
            i -= 1
        assert col(out, "tp")[r] == np.float32(i * 16.0 + raw_tbl["waveform"].t0.nda[r])


def test_round_floor_ceil_trunc(raw_tbl):
    """:393-449"""
    cfg = {"outputs": ["w_round", "w_floor", "w_ceil", "w_trunc"], "processors": {
        "w_round": "round(waveform, 4)", "w_floor": "floor(waveform, 4)", "w_ceil": "ceil(waveform, 4)",
        "w_trunc": "trunc(waveform, 4)"}}
    out = build(raw_tbl, cfg, n_entries=1)
    wf = raw_tbl["waveform"].values.nda[0]
    assert np.all(np.rint(wf / 4) * 4 == wvals(out, "w_round")[0])
    assert np.all(np.floor(wf / 4) * 4 == wvals(out, "w_floor")[0])
    assert np.all(np.ceil(wf / 4) * 4 == wvals(out, "w_ceil")[0])
    assert np.all(np.trunc(wf / 4) * 4 == wvals(out, "w_trunc")[0])
    cfg = {"outputs": ["tp_max", "t_round", "t_floor", "t_ceil", "t_trunc", "c_round", "c_floor", "c_ceil", "c_trunc"],
           "processors": {
               "tp_min, tp_max, wf_min, wf_max": {"function": "min_max", "module": "dspeed.processors",
                                                  "args": ["waveform", "tp_min", "tp_max", "wf_min", "wf_max"],
                                                  "unit": ["us", "us", "ADC", "ADC"]},
               "t_round": "round(tp_max, 1*us)", "t_floor": "floor(tp_max, 1*us)", "t_ceil": "ceil(tp_max, 1*us)",
               "t_trunc": "trunc(tp_max, 1*us)",
               "c_round": "round(1*us, waveform.period)", "c_floor": "floor(1*us, waveform.period)",
               "c_ceil": "ceil(1*us, waveform.period)", "c_trunc": "trunc(1*us, waveform.period)"}}
    out = build(raw_tbl, cfg, n_entries=1)
    tp = col(out, "tp_max")[0]
    assert tp == np.float32(np.argmax(raw_tbl["waveform"].values.nda[0]) * 0.016)
    assert np.rint(tp) == col(out, "t_round")[0] and np.floor(tp) == col(out, "t_floor")[0]
    assert np.ceil(tp) == col(out, "t_ceil")[0] and np.trunc(tp) == col(out, "t_trunc")[0]
    assert col(out, "c_round")[0] == 992      # 62.5 periods: round half to even
    assert col(out, "c_floor")[0] == 992 and col(out, "c_ceil")[0] == 1008 and col(out, "c_trunc")[0] == 992


def test_where_and_ternary(raw_tbl):
    """:452-587"""
    from dspeed_b200.errors import ProcessingChainError

    mm = {"function": "min_max", "module": "dspeed.processors", "args": ["waveform", "tp_min", "tp_max", "wf_min", "wf_max"],
          "unit": ["ns", "ns", "ADC", "ADC"]}
    cfg = {"outputs": ["tp_min", "tp_max", "wf_min", "wf_max", "test1", "test2", "test3", "test4", "test5", "test6"],
           "processors": {
               "tp_min, tp_max, wf_min, wf_max": mm,
               "test1": "where(waveform<12100, 0, waveform)", "test2": "where(waveform<12100, waveform, 0)",
               "test3": "where(eventnumber==0, tp_min, 1*ns)", "test4": "where(eventnumber==0, tp_min, 1*us)",
               "test5": "where(eventnumber==0, 1*ns, tp_min)", "test6": "where(eventnumber==0, 1*us, tp_min)",
               "test7": "where(eventnumber==0, tp_min, wf_min)"}}
    out = build(raw_tbl, cfg, n_entries=2)
    wf = raw_tbl["waveform"].values.nda[0]
    assert np.all(np.where(wf < 12100, 0, wf) == wvals(out, "test1")[0])
    assert np.all(np.where(wf < 12100, wf, 0) == wvals(out, "test2")[0])
    tp_min = col(out, "tp_min")
    # a time constant chosen against a coordinate variable becomes a coordinate on that variable's grid (reference
    # :1392-1409: const / period, is_coord of the variable), so on output it is shifted by the event's t0 like the
    # variable itself -- the reference test sees the bare constants only because its file has t0 = 0; here row 1 has
    # t0 = 32 ns
    for name, first, second in (("test3", tp_min[0], 1 + 32), ("test4", tp_min[0], 1000 + 32), ("test5", 1, tp_min[1]),
                                ("test6", 1000, tp_min[1])):
        assert out[name].attrs["units"] == "ns", name
        assert col(out, name)[0] == first and col(out, name)[1] == second, name
    with pytest.raises(ProcessingChainError):          # a time and an amplitude have no common unit
        build(raw_tbl, cfg, outputs=["test7"])
    cfg = {"processors": {
        "w_downsample": "waveform[::2]", "w_win1": "waveform[:len(waveform)//2]", "w_win2": "waveform[len(waveform)//2:]",
        "tp_min, tp_max, wf_min, wf_max": mm, "delta_t": "tp_max - tp_min",
        "test1": "where(eventnumber==0, w_downsample, w_win1)",   # different periods
        "test2": "where(eventnumber==0, w_win1, w_win2)",         # different offsets
        "test3": "where(eventnumber==0, tp_max, delta_t)",        # a coordinate and a duration
        "test4": "where(eventnumber==0, tp_min, tp_max)"}}
    with pytest.raises(ProcessingChainError):
        build(raw_tbl, cfg, outputs=["test1"])
    out = build(raw_tbl, cfg, outputs=["test2", "w_win1", "w_win2"], n_entries=2)
    t0 = lambda name: np.asarray(out[name].t0.nda.cpu() if hasattr(out[name].t0.nda, "cpu") else out[name].t0.nda)  # noqa: E731
    assert np.all(wvals(out, "test2")[0] == wvals(out, "w_win1")[0]) and t0("test2")[0] == t0("w_win1")[0]
    assert np.all(wvals(out, "test2")[1] == wvals(out, "w_win2")[1]) and t0("test2")[1] == t0("w_win2")[1]
    assert t0("w_win2")[1] == 32.0 + 4000 * 16.0
    with pytest.raises(ProcessingChainError):
        build(raw_tbl, cfg, outputs=["test3"])
    out = build(raw_tbl, cfg, outputs=["test4", "tp_min", "tp_max"], n_entries=2)
    assert out["test4"].attrs["units"] == "ns"
    assert col(out, "test4")[0] == col(out, "tp_min")[0] and col(out, "test4")[1] == col(out, "tp_max")[1]
    cfg = {"processors": {
        "test1": "where(eventnumber==0, 10*ns, 1*us, dtype='f')", "test2": "where(eventnumber==0, 10*ns, 1000, dtype='f')",
        "test3": "where(eventnumber==0, 1000, 10*ns, dtype='f')", "test4": "where(eventnumber==0, 10, 1000, dtype='f')",
        "test5": "where(eventnumber==0, 10*ns, 10*m, dtype='f')"}}
    out = build(raw_tbl, cfg, outputs=["test1", "test2", "test3", "test4"], n_entries=2)
    for name, a, b in (("test1", 10, 1000), ("test2", 10, 1000), ("test3", 1000, 10), ("test4", 10, 1000)):
        assert col(out, name)[0] == a and col(out, name)[1] == b, name
        if name != "test4":
            assert out[name].attrs["units"] == "ns"
    with pytest.raises(ProcessingChainError):
        build(raw_tbl, cfg, outputs=["test5"])
    out = build(raw_tbl, {"outputs": ["test"], "processors": {"test": "0 if waveform<12100 else waveform"}}, n_entries=1)
    assert np.all(np.where(wf < 12100, 0, wf) == wvals(out, "test")[0])


def test_isnan_isfinite_astype(raw_tbl):
    """:590-620"""
    from dspeed_b200 import tables

    tb = tables.Table({"input": tables.Array(np.array([1.0, 0.0, np.inf, -np.inf, np.nan]))}, size=5)
    out = build(tb, {"outputs": ["test_nan", "test_finite"], "processors": {"test_nan": "isnan(input)", "test_finite": "isfinite(input)"}})
    assert np.all(np.array([False, False, False, False, True]) == col(out, "test_nan"))
    assert np.all(np.array([True, True, False, False, False]) == col(out, "test_finite"))
    out = build(raw_tbl, {"outputs": ["waveform_32"], "processors": {"waveform_32": "astype(waveform, 'float32')"}}, n_entries=1)
    assert wvals(out, "waveform_32").dtype == np.float32
    assert np.all(raw_tbl["waveform"].values.nda[0] == wvals(out, "waveform_32")[0])


def test_mean_stdev_alias(raw_tbl):
    """BASELINE.json names the baseline processor `mean_stdev`: the first two outputs of linear_slope_fit
    (linear_slope_fit.py:11-90), here against float64 arithmetic on the same samples"""
    cfg = {"outputs": ["bl_mean", "bl_std", "m2", "s2"], "processors": {
        "bl_mean, bl_std": {"function": "mean_stdev", "module": "dspeed.processors", "args": ["waveform[0:750]", "bl_mean", "bl_std"],
                            "unit": ["ADC", "ADC"]},
        "m2, s2, sl, ic": {"function": "linear_slope_fit", "module": "dspeed.processors",
                           "args": ["waveform[0:750]", "m2", "s2", "sl", "ic"], "unit": ["ADC"] * 4}}}
    out = build(raw_tbl, cfg)
    x = raw_tbl["waveform"].values.nda[:, :750].astype(np.float64)
    assert np.array_equal(col(out, "bl_mean"), col(out, "m2")) and np.array_equal(col(out, "bl_std"), col(out, "s2"))
    assert np.abs(col(out, "bl_mean") - x.mean(1)).max() <= 2e-7 * x.mean(1).max()
    # closer to the float64 truth than float32 rounding of the result itself allows to tell
    assert np.abs(col(out, "bl_std") - x.std(1, ddof=1)).max() <= 4e-7 * x.std(1, ddof=1).max()
