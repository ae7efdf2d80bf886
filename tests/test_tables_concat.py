"""build_dsp over a multi-chunk in-memory iterator appends the chunks' output tables column by column, whatever the
column kind (the reference appends through lh5 / lgdo generically, build_dsp.py:416-424)."""
import numpy as np
import torch

from dspeed_b200 import tables
from dspeed_b200.build_dsp import _concat_tables


def _part(n, seed):
    r = np.random.default_rng(seed)
    cl = np.cumsum(r.integers(0, 4, n)).astype(np.uint32)
    return tables.Table({
        "a": tables.Array(r.normal(size=n).astype(np.float32), attrs={"units": "ADC"}),
        "m": tables.ArrayOfEqualSizedArrays(r.normal(size=(n, 3))),
        "w": tables.WaveformTable(size=n, t0=tables.Array(r.normal(size=n), attrs={"units": "ns"}), dt=16, dt_units="ns",
                                  values=r.normal(size=(n, 5)).astype(np.float32)),
        "v": tables.VectorOfVectors(flattened_data=r.normal(size=int(cl[-1])), cumulative_length=cl),
        "d": tables.Array(torch.arange(n, dtype=torch.float32)),       # a tensor column stays a tensor
    }, size=n)


def test_chunks_of_every_column_kind_are_appended():
    p1, p2, p3 = _part(4, 1), _part(3, 2), _part(5, 3)
    o = _concat_tables([p1, p2, p3])
    assert len(o) == 12 and o["a"].attrs["units"] == "ADC"
    assert np.array_equal(o["a"].nda, np.concatenate([p["a"].nda for p in (p1, p2, p3)]))
    assert isinstance(o["m"], tables.ArrayOfEqualSizedArrays) and o["m"].nda.shape == (12, 3)
    assert tables.kind_of(o["w"]) == "wftable" and o["w"].values.nda.shape == (12, 5)
    assert o["w"].t0.attrs["units"] == "ns" and o["w"].dt.attrs["units"] == "ns"
    assert np.array_equal(o["w"].t0.nda, np.concatenate([p["w"].t0.nda for p in (p1, p2, p3)]))
    assert isinstance(o["d"].nda, torch.Tensor) and o["d"].nda.shape == (12,)
    rows = [(p, i) for p in (p1, p2, p3) for i in range(len(p))]
    for k, (p, i) in enumerate(rows):
        assert np.array_equal(o["v"][k], p["v"][i]), k
    assert int(o["v"].cumulative_length.nda[-1]) == sum(int(p["v"].cumulative_length.nda[-1]) for p in (p1, p2, p3))
