"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes shard the rows, and
the single collective of the path (the gather of the output tables) reassembles them in
order.  No CUDA work happens here; the per-shard chain itself is covered by the -m gpu tests."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dspeed_b200 import parallel, tables


def test_shard_rows_tile_the_table():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [parallel.shard_rows(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_rows(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = parallel.shard_rows(n_rows, rank, world)
        rows = np.arange(b, e)
        # what a rank's chain would produce for its shard: values that identify the global row
        local = tables.Table({
            "trapEmax": tables.Array((rows * 2.0).astype(np.float32), attrs={"units": "ADC"}),
            "tp_0_est": tables.Array((rows + 0.5).astype(np.float32), attrs={"units": "ns"}),
            "wf": tables.ArrayOfEqualSizedArrays(np.repeat(rows[:, None], 3, axis=1).astype(np.float32)),
            # the SiPM chain's outputs: NaN-padded peak lists [rows, 20] and uint32 counters
            "vt_max": tables.ArrayOfEqualSizedArrays(np.where(np.arange(20)[None, :] < (rows % 5)[:, None],
                                                              rows[:, None] + np.arange(20)[None, :], np.nan).astype(np.float32)),
            "n_max": tables.Array((rows % 5).astype(np.uint32)),
            # variable-length (VectorOfVectors) column: row r holds r % 4 entries r, r + 0.5, ...
            "trigger_pos": tables.VectorOfVectors(
                flattened_data=np.concatenate([r + 0.5 * np.arange(r % 4) for r in rows] + [np.zeros(0)]).astype(np.float64),
                cumulative_length=np.cumsum(rows % 4).astype(np.uint32), attrs={"units": "ns"}),
            # waveform output with a per-event t0
            "wf_out": tables.WaveformTable(size=e - b, t0=tables.Array(rows * 16.0, attrs={"units": "ns"}), dt=16, dt_units="ns",
                                           values=np.repeat(rows[:, None], 4, axis=1).astype(np.float32)),
        }, size=e - b)
        full = parallel.gather_table(local, n_rows, dst=0)
        if rank == 0:
            ok = (len(full) == n_rows
                  and np.array_equal(np.asarray(full["trapEmax"].nda), np.arange(n_rows, dtype=np.float32) * 2)
                  and np.array_equal(np.asarray(full["tp_0_est"].nda), np.arange(n_rows, dtype=np.float32) + 0.5)
                  and np.array_equal(np.asarray(full["wf"].nda)[:, 2], np.arange(n_rows, dtype=np.float32))
                  and full["trapEmax"].attrs["units"] == "ADC"
                  and np.asarray(full["n_max"].nda).dtype == np.uint32
                  and np.array_equal(np.asarray(full["n_max"].nda), (np.arange(n_rows) % 5).astype(np.uint32))
                  and np.array_equal(np.isnan(np.asarray(full["vt_max"].nda)).sum(axis=1), 20 - np.arange(n_rows) % 5)
                  and np.array_equal(np.asarray(full["vt_max"].nda)[1::5, 0], np.arange(n_rows, dtype=np.float32)[1::5])
                  and np.array_equal(np.asarray(full["trigger_pos"].cumulative_length.nda), np.cumsum(np.arange(n_rows) % 4))
                  and all(np.array_equal(full["trigger_pos"][r], r + 0.5 * np.arange(r % 4)) for r in range(n_rows))
                  and full["trigger_pos"].attrs["units"] == "ns"
                  and np.array_equal(np.asarray(full["wf_out"].t0.nda), np.arange(n_rows) * 16.0)
                  and np.array_equal(np.asarray(full["wf_out"].values.nda)[:, 3], np.arange(n_rows, dtype=np.float32))
                  and float(np.asarray(full["wf_out"].dt.nda)[0]) == 16.0)
            q.put(bool(ok))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [1, 11, 64])
def test_gather_of_output_tables_gloo_world2(n_rows):
    """one packed gather for every fixed-shape column (+ one for the ragged payloads); n_rows = 1: rank 1's shard is
    empty"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
