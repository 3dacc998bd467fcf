"""The device-resident form of the ABI (rl_batch_*): upload validation, re-upload of new values, the rows a solve may
touch, device-side views of the outputs and the final gather over NCCL."""
import os
import sys

import numpy as np
import pytest

import practice_path_planning_for_formula_student_driverless_b200 as rl
from conftest import ROOT, TOL_LAP_REL, assert_result_close
from test_gpu_parity import MC, MT, oracle_ref, stalled

pytestmark = pytest.mark.gpu


def _tracks(n_tracks, n, seed):
    center, seg, L, m = rl.synth_tracks(n_tracks, n, seed_base=seed)
    center, seg = center.reshape(n_tracks, n, 2), seg.reshape(n_tracks, 2, m, 4)
    return [rl.Track(center[i].copy(), seg[i, 0].copy(), seg[i, 1].copy(), float(L[i])) for i in range(n_tracks)]


def test_reupload_new_values_then_solve(ctx):
    """rl_batch_upload with the same shapes but new geometry, new params and another job -> param mapping."""
    a, b = _tracks(3, 300, 0xA110), _tracks(3, 300, 0xA220)
    cfgs_a = [rl.Config(), rl.Config(lambda_smooth=3.2e-3)]
    cfgs_b = [rl.Config(w_time_gain=2.5, P_max_W=40000.0), rl.Config(safety_margin_m=0.15)]
    jobs_a = [(0, 0, MC), (0, 0, MT), (1, 1, MT), (2, 0, MC)]
    jobs_b = [(0, 1, MC), (0, 0, MT), (1, 0, MT), (2, 1, MC)]       # same (track, stage), other params
    pa = rl.PackedBatch(a, [c.to_params() for c in cfgs_a], jobs_a)
    dev = rl.DeviceBatch(ctx, pa)
    dev.solve(); dev.download(); dev.sync()
    for j, (t, c, st) in enumerate(jobs_a):
        o = oracle_ref(st, a[t], cfgs_a[c].to_params())
        assert_result_close(pa.result(j), o, "o_", st == MT, tag=("first upload", j))
    pb = rl.PackedBatch(b, [c.to_params() for c in cfgs_b], jobs_b)
    dev.host = pb
    dev.upload(); dev.solve(); dev.download(); dev.sync()
    for j, (t, c, st) in enumerate(jobs_b):
        o = oracle_ref(st, b[t], cfgs_b[c].to_params())
        r = pb.result(j)
        assert_result_close(r, o, "o_", st == MT, tag=("re-upload", j))
        assert r.stats.accepted == o["stats"].accepted
        if not stalled(o["stats"]):
            assert r.stats.backtracks == o["stats"].backtracks
        if st == MT:
            assert abs(r.lap_time - o["lap"]) <= TOL_LAP_REL * o["lap"]
    dev.close()


def test_upload_rejects_other_shapes(ctx):
    """The plan is frozen at rl_batch_create: same totals but other per-track sizes, closed flags or job -> track / stage
    mappings must be refused (they would run kernels of the wrong size class)."""
    t300, t200, t400 = _tracks(2, 300, 0xB1), _tracks(1, 200, 0xB2), _tracks(1, 400, 0xB3)
    jobs = [(0, 0, MC), (1, 0, MT)]
    base = rl.PackedBatch(t300, [rl.Config().to_params()], jobs)
    dev = rl.DeviceBatch(ctx, base)

    def refused(batch):
        dev.host = batch
        with pytest.raises(rl.RacelineError) as e:
            dev.upload()
        assert e.value.status == rl.RL_ERR_ARG

    # same total samples (600) and the same total segments, other split
    m = t300[0].inner_seg.shape[0]
    other = [rl.Track(t200[0].center_xy, t300[0].inner_seg, t300[0].outer_seg, t200[0].L),
             rl.Track(t400[0].center_xy, t300[1].inner_seg, t300[1].outer_seg, t400[0].L)]
    assert other[0].inner_seg.shape[0] == m
    refused(rl.PackedBatch(other, [rl.Config().to_params()], jobs))
    refused(rl.PackedBatch(t300, [rl.Config().to_params()], [(1, 0, MC), (0, 0, MT)]))     # job -> track
    refused(rl.PackedBatch(t300, [rl.Config().to_params()], [(0, 0, MT), (1, 0, MT)]))     # job -> stage
    opened = [rl.Track(t.center_xy, t.inner_seg, t.outer_seg, t.L, closed=(i == 0)) for i, t in enumerate(t300)]
    refused(rl.PackedBatch(opened, [rl.Config().to_params()], jobs))                        # closed flag
    dev.host = base
    dev.upload(); dev.solve(); dev.download(); dev.sync()                                   # the batch is still usable
    assert base.out_stats[0].status == 0 and base.out_stats[1].lap_time > 0
    dev.close()


@pytest.mark.parametrize("layout", ["interleaved", "irregular"])
def test_v_and_ax_rows_of_mincurv_jobs_are_left_untouched(ctx, layout):
    """include/raceline_b200.h: 'v/ax rows of MINCURV jobs are left untouched' -- through rl_solve_batch (chunked,
    strided copy for the regular interleave) and through rl_batch_download."""
    tr = _tracks(4, 256, 0xC0DE)
    if layout == "interleaved":
        jobs = [(t, 0, st) for t in range(4) for st in (MC, MT)]
    else:
        jobs = [(0, 0, MT), (1, 0, MC), (1, 0, MT), (2, 0, MT), (3, 0, MC), (0, 0, MC), (3, 0, MT)]
    for resident in (False, True):
        pb = rl.PackedBatch(tr, [rl.Config().to_params()], jobs)
        pb.out_v[:] = -7.0; pb.out_ax[:] = -9.0
        if resident:
            dev = rl.DeviceBatch(ctx, pb)
            dev.solve(); dev.download(); dev.sync(); dev.close()
        else:
            ctx.solve_batch(pb)
        for j, (_, _, st) in enumerate(jobs):
            a, b = int(pb.job_off[j]), int(pb.job_off[j + 1])
            if st == MC:
                assert np.all(pb.out_v[a:b] == -7.0) and np.all(pb.out_ax[a:b] == -9.0), (layout, resident, j)
            else:
                assert np.all(pb.out_v[a:b] > 0.0) and np.all(pb.out_ax[a:b] != -9.0), (layout, resident, j)


def test_many_pipeline_chunks_equal_one(ctx):
    """rl_solve_batch with 1 and with 16 pipeline chunks, the 16 spread over the kernel streams in each of the three
    ways rl_set_option("chunk_streams") knows: bit-identical results (chunks only cut the job list)."""
    tr = _tracks(40, 128, 0xD00D)
    jobs = [(t, 0, st) for t in range(40) for st in (MC, MT)]
    outs = []
    for chunks, streams in ((1, 0), (16, 0), (16, 1), (16, 2), (16, 3)):
        ctx.set_option("solve_chunks", chunks)
        ctx.set_option("chunk_streams", streams)
        pb = rl.PackedBatch(tr, [rl.Config().to_params()], jobs)
        ctx.solve_batch(pb)
        outs.append(pb)
    ctx.set_option("solve_chunks", 0)
    ctx.set_option("chunk_streams", 0)
    a = outs[0]
    for b in outs[1:]:
        assert np.array_equal(a.out_xy, b.out_xy) and np.array_equal(a.out_alpha_total, b.out_alpha_total)
        for j, (_, _, st) in enumerate(jobs):
            if st == MT:
                assert a.out_stats[j].lap_time == b.out_stats[j].lap_time
                lo, hi = int(a.job_off[j]), int(a.job_off[j + 1])
                assert np.array_equal(a.out_v[lo:hi], b.out_v[lo:hi])
    with pytest.raises(rl.RacelineError):
        ctx.set_option("no_such_option", 1)


def test_device_tensors_alias_the_outputs(ctx):
    import torch
    tr = _tracks(3, 192, 0xE1)
    jobs = [(t, 0, st) for t in range(3) for st in (MC, MT)]
    pb = rl.PackedBatch(tr, [rl.Config().to_params()], jobs)
    dev = rl.DeviceBatch(ctx, pb)
    dev.solve(); dev.download(); dev.sync()
    laps = dev.device_tensor("lap_time")
    assert laps.is_cuda and laps.shape == (6,)
    assert np.array_equal(laps.cpu().numpy(), np.array([pb.out_stats[j].lap_time for j in range(6)]))
    assert np.array_equal(dev.device_tensor("xy").cpu().numpy(), pb.out_xy)
    assert np.array_equal(dev.device_tensor("curvature").cpu().numpy(), pb.out_curvature)
    assert torch.count_nonzero(laps[0::2]).item() == 0 and torch.all(laps[1::2] > 0)
    dev.close()


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from practice_path_planning_for_formula_student_driverless_b200 import sharding
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    # a sharded Config sweep on one synthetic track: 9 problems over 2 ranks
    n, total = 160, 9
    tr = _tracks(1, n, 0xF00D)
    cfgs = [rl.Config(w_time_gain=0.4 * k, lambda_smooth=1.6e-3 * (1 + 0.2 * k)) for k in range(total)]
    lo, hi = sharding.shard_bounds(total, world, rank)
    jobs = [(0, k, MT) for k in range(lo, hi)]
    ctx = rl.Context(rank)
    pb = rl.PackedBatch(tr, [c.to_params() for c in cfgs], jobs)
    dev = rl.DeviceBatch(ctx, pb)
    dev.solve(); dev.sync()
    laps = sharding.gather_lap_times(dev.device_tensor("lap_time"), total)           # device tensors, NCCL
    xy = sharding.gather_rasters(dev.device_tensor("xy"), total, n).cpu().numpy()
    best = sharding.best_of_sweep(dev.device_tensor("lap_time"), total)
    dev.close(); ctx.close()
    q.put((rank, laps, xy, best))
    dist.barrier()
    dist.destroy_process_group()


def test_final_gather_over_nccl_two_ranks(ctx):
    """north_star: 'only a final gather of per-problem lap times and rasters (NCCL over NVLink)'.  Two ranks solve the
    halves of a sweep; both must end with exactly what a single rank computes for the whole sweep."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    n, total = 160, 9
    tr = _tracks(1, n, 0xF00D)
    cfgs = [rl.Config(w_time_gain=0.4 * k, lambda_smooth=1.6e-3 * (1 + 0.2 * k)) for k in range(total)]
    single = rl.solve_batch(tr, cfgs, [(0, k, MT) for k in range(total)], ctx=ctx)
    expect_laps = np.array([r.lap_time for r in single])
    expect_xy = np.concatenate([r.raceline for r in single])
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_nccl_worker, args=(r, 2, 29741, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for _, laps, xy, best in got:
        assert np.array_equal(laps, expect_laps)
        assert np.array_equal(xy, expect_xy)
        assert best[0] == int(np.argmin(expect_laps)) and best[1] == expect_laps.min()
