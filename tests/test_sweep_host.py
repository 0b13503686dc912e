"""Host-side logic of the multi-GPU sweep (pcm/sweep.py): grid construction, sharding and the
single final gather -- exercised with world_size 2 over gloo on CPU.  The per-sequence GPU work
itself is covered by tests/test_gpu_sequence.py."""
import os
import socket

import numpy as np
import pytest


def test_grid_matches_reference_benchmark_py():
    from pcm import sweep
    plist = sweep.params_list()
    assert len(plist) == 64                                   # benchmark.py:42-51: 2*2*1*2*2*2*1*2
    assert plist[0] == dict(n_estimators=20, max_depth=7, n_components=1, novelty_detection=True,
                            over_segmentation="quickshift", features="8 hsv_lab", dilation_kernel=7, prior_weight=0.1)
    # itertools.product order: the last key varies fastest
    assert plist[1]["prior_weight"] == 0.0 and plist[2]["features"] == "6 lab"
    items = sweep.build_items()
    assert len(items) == 256
    assert [it[0] for it in items] == list(range(256))
    assert items[5][1:3] == (1, "frog")                       # id = i * len(videos) + k


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8, 40])
def test_partition_is_a_balanced_exact_cover(world):
    from pcm import sweep
    items = sweep.build_items()
    shards = sweep.partition(items, world)
    assert len(shards) == world
    ids = sorted(it[0] for s in shards for it in s)
    assert ids == list(range(256))                            # every sequence exactly once
    assert all(len(s) > 0 for s in shards)
    loads = [sweep.shard_cost(s, items) for s in shards]       # sequences + their clips' preparation + clip opening
    if world <= 8:
        assert max(loads) <= 1.12 * (sum(loads) / world)
        seq_loads = [sum(sweep.item_cost(it) for it in s) for s in shards]
        assert max(seq_loads) <= 1.25 * (sum(seq_loads) / world)
        # clip locality: a rank decodes / over-segments / runs SIFT on the clips it touches -- at most two, and
        # three at the end of the list where the two short clips sit
        n_clips = [len({it[2] for it in s}) for s in shards]
        assert world == 1 or (all(n <= 3 for n in n_clips) and (world < 4 or sorted(n_clips)[len(n_clips) // 2] <= 2))
    assert sweep.partition(items, world) == shards            # deterministic


def test_results_table_has_the_reference_columns():
    from pcm import sweep
    videos = ["soldier", "frog"]
    hyper = dict(n_estimators=[20, 30], max_depth=[5], prior_weight=[0.0])
    rows = np.array([[0, 0.5, 1.0], [1, 0.7, 2.0], [2, 0.6, 1.5], [3, 0.8, 2.5]])
    t = sweep.results_table(rows, videos, hyper)
    assert list(t.columns) == ["n_estimators", "max_depth", "prior_weight", "soldier_benchmark", "soldier_time",
                               "frog_benchmark", "frog_time", "avg_benchmark"]
    assert t.loc[0, "avg_benchmark"] == pytest.approx(0.6) and t.loc[1, "frog_time"] == 2.5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from pcm import sweep
    dist.init_process_group("gloo", rank=rank, world_size=world)
    items = sweep.build_items(["soldier", "bmx"], dict(n_estimators=[20, 30], max_depth=[7, 10], features=["6 lab"], prior_weight=[0.0, 0.1]))
    shard = sweep.partition(items, world)[rank]
    # stand-in for run_shard: a deterministic score per sequence id
    local = np.array([[it[0], 0.25 + 0.01 * it[0], 1.0 + it[0]] for it in shard], np.float64).reshape(-1, 3)
    allr = sweep.gather(local, world, dist, None)
    q.put((rank, len(shard), allr))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_items = 2 * 8
    assert sum(g[1] for g in got) == n_items and all(g[1] > 0 for g in got)
    for _, _, allr in got:                                    # every rank holds the full, sorted table
        assert allr.shape == (n_items, 3)
        assert np.array_equal(allr[:, 0], np.arange(n_items))
        assert np.allclose(allr[:, 1], 0.25 + 0.01 * np.arange(n_items))
        assert np.allclose(allr[:, 2], 1.0 + np.arange(n_items))


def test_gather_single_rank_needs_no_collective():
    from pcm import sweep
    local = np.array([[3, 0.1, 1.0], [1, 0.2, 2.0]])
    out = sweep.gather(local, 1)
    assert np.array_equal(out[:, 0], [1, 3])


def test_smaller_forest_is_a_prefix_of_the_larger_one():
    """What `fit_estimators` relies on: with random_state=42 (the reference's, :199) tree i of a
    20-tree forest IS tree i of the 30-tree forest of the same depth on the same rows."""
    from sklearn.ensemble import RandomForestClassifier
    rng = np.random.default_rng(3)
    X = rng.integers(-1, 256, (3000, 60)).astype(np.float64) / 255
    y = ((X[:, 5] > 0.5) ^ (X[:, 17] > 0.3)).astype(np.int64)
    small = RandomForestClassifier(random_state=42, n_estimators=20, max_depth=7).fit(X, y)
    large = RandomForestClassifier(random_state=42, n_estimators=30, max_depth=7, n_jobs=3).fit(X, y)
    for a, b in zip(small.estimators_, large.estimators_[:20]):
        ta, tb = a.tree_, b.tree_
        assert np.array_equal(ta.feature, tb.feature) and np.array_equal(ta.threshold, tb.threshold)
        assert np.array_equal(ta.children_left, tb.children_left) and np.array_equal(ta.value, tb.value)
    # and the probability of the small forest is the mean over that prefix
    p = np.mean([e.predict_proba(X)[:, 1] for e in large.estimators_[:20]], axis=0)
    np.testing.assert_allclose(p, small.predict_proba(X)[:, 1], rtol=0, atol=1e-15)


def test_model_cache_computes_once_under_threads():
    import threading
    import time
    from pcm.sweep import ModelCache
    cache, calls = ModelCache(), []

    def make():
        calls.append(1)
        time.sleep(0.05)
        return object()
    out = []
    ts = [threading.Thread(target=lambda: out.append(cache.get_or_compute(("k", 1), make))) for _ in range(8)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert len(calls) == 1 and all(o is out[0] for o in out)
    assert cache.get_or_compute(("k", 2), make) is not out[0] and len(calls) == 2


def _run_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from pcm import sequence, sweep

    def fake_run_sequence(cfg, device=0, model_cache=None, cache_tag=None, max_frames=None, **kw):
        # stands in for the GPU work of one sequence: a score that depends on the config only
        p = cfg["params"]
        iou = 0.5 + 0.001 * p["n_estimators"] + 0.01 * p["max_depth"] + (0.05 if "frog" in cfg["input_video"] else 0.0)
        return dict(mean_iou=iou, seconds=0.25, train_seconds=0.0, decode_seconds=0.0, wall_seconds=0.25)
    sequence.run_sequence = fake_run_sequence
    hyper = dict(n_estimators=[20, 30], max_depth=[7, 10], features=["6 lab"], prior_weight=[0.0])
    summary, table = sweep.run({"masker": "PC"}, {"soldier": {}, "frog": {}}, videos=["soldier", "frog"], hyper=hyper,
                               backend="gloo", seq_workers=2, resident=False)
    q.put((rank, summary, None if table is None else table.to_dict()))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_run_world_size_2_gloo_end_to_end():
    """sweep.run() with two ranks over gloo (sequence work stubbed out): sharding, the sequence
    threads, the single gather, the max-over-ranks time and rank 0's results table."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_run_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict((g[0], g[1:]) for g in [q.get(timeout=180) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s0, t0 = got[0]
    s1, t1 = got[1]
    assert t1 is None and t0 is not None                      # the table lives on rank 0
    assert s0["n_sequences"] == 8 and s0["n_gpus"] == 2 and sum(s0["per_rank_sequences"]) == 8
    assert s0["seconds"] == s1["seconds"] > 0                 # max over ranks, the same on both
    assert t0["soldier_benchmark"][0] == pytest.approx(0.5 + 0.02 + 0.07)
    assert t0["frog_benchmark"][3] == pytest.approx(0.5 + 0.03 + 0.10 + 0.05)
    assert t0["avg_benchmark"][0] == pytest.approx((0.59 + 0.64) / 2)
    assert all(v == 0.25 for v in t0["frog_time"].values())


def _share_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from pcm.fastseq import Share
    dist.init_process_group("gloo", rank=rank, world_size=world)
    group = dist.new_group(ranks=list(range(world)))
    share = Share(dist, group, rank, world)
    n = 11
    mine = list(share.mine(n))
    flat = torch.zeros(n * 3, dtype=torch.int32)
    desc = torch.zeros((n, 4), dtype=torch.uint8)
    for k in mine:                                   # every member fills only its own entries ...
        flat[3 * k:3 * k + 3] = torch.tensor([k, 10 * k, 100 + k], dtype=torch.int32)
        desc[k] = torch.tensor([k, 255 - k, 7, 0], dtype=torch.uint8)
    share.combine(flat)                              # ... and the all-reduce brings in the others'
    share.combine(desc)
    q.put((rank, mine, flat.numpy().copy(), desc.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_clip_share_splits_and_combines_world_size_2_gloo():
    """fastseq.Share: the ranks working on one clip take interleaved entries of its per-frame host work and combine the
    pieces with one all-reduce per buffer (NCCL on GPUs; gloo here)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_share_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in procs], key=lambda g: g[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == [0, 2, 4, 6, 8, 10] and got[1][1] == [1, 3, 5, 7, 9]
    want = np.array([[k, 10 * k, 100 + k] for k in range(11)], np.int32).reshape(-1)
    for _, _, flat, desc in got:
        assert np.array_equal(flat, want)
        assert np.array_equal(desc[:, 0], np.arange(11)) and np.array_equal(desc[:, 1], 255 - np.arange(11))


class _FakeTensor:
    def __init__(self, address):
        self.address = address

    def data_ptr(self):
        return self.address


def _bare_masker(frame_numbers, multi_selection, index=0, current_model=0, novelty=True, prior_weight=0.1):
    from maskers.pixel_classification import PixelClassificationNonRigidMasker
    m = object.__new__(PixelClassificationNonRigidMasker)
    m.config = dict(params=dict(novelty_detection=novelty, dilation_kernel=7, prior_weight=prior_weight))
    m.multi_selection, m.index, m.current_model = multi_selection, index, current_model
    m.models = [dict(n_frame=f) for f in frame_numbers]
    m.novelty_det = [dict(n_frame=f, threshold=0.25 + 0.125 * k) for k, f in enumerate(frame_numbers)]
    m.prevFrame = m.prevForegroundMask = None
    return m


@pytest.mark.parametrize("frame_numbers,multi,n,index,with_prior,n_truth", [
    ([0, 93, 186], True, 279, 0, True, 279),          # frog: two blended spans, the third model alone
    ([0, 93, 186], True, 40, 90, False, 12),          # starts inside a span, crosses one switch
    ([0, 10], False, 32, 0, True, 0),                 # multi_selection off: the first model throughout
    ([0], True, 1, 0, True, 5),                       # a single frame
    ([0, 5, 7], True, 20, 6, False, 20),              # already past the next selection: switches after one frame
])
def test_build_jobs_equals_the_per_frame_fill(frame_numbers, multi, n, index, with_prior, n_truth):
    """fastseq.build_jobs (column-wise numpy fill of the pcm_frame_job records) against the per-frame ctypes fill
    driven by the masker's own _frame_params() / _advance() state machine: same bytes, same final state."""
    from pcm import capi, fastseq
    rng = np.random.RandomState(3)
    H, W, fb = 240, 320, 240 * 320 * 3
    rects = [tuple(int(v) for v in rng.randint(0, 100, 4)) for _ in range(n)]
    sizes = rng.randint(50, 500, n)
    arena = fastseq.LabelArena(_FakeTensor(0x7f0000001000), np.concatenate([[0], np.cumsum(sizes)]).tolist(),
                               [int(v) for v in rng.randint(2, 90, n)])
    counts = rng.randint(0, 40, n)
    sift = fastseq.SiftStore(None, None, _FakeTensor(0x7f1000000000), _FakeTensor(0x7f2000000000),
                             np.concatenate([[0], np.cumsum(counts)]).tolist()) if with_prior else None
    frames_ptr, pri_ptr, truth_ptr, counts_ptr = 0x7f3000000000, 0x7f4000000100 if with_prior else 0, 0x7f5000000000, 0x7f6000000000

    a = _bare_masker(frame_numbers, multi, index=index)
    jobs = (capi.FrameJob * n)()
    for k in range(n):
        j = jobs[k]
        j.d_frame = frames_ptr + k * fb
        j.rect[:] = rects[k]
        j.d_labels = arena.ptr(k)
        j.n_labels = arena.n_labels[k]
        j.clear_mask = 2 if k > 0 else 0
        p, blend = a._frame_params()
        j.params = p
        if with_prior and k > 0:
            j.d_pts_prev, j.d_des_prev, j.n_prev = sift.pts_ptr(k - 1), sift.des_ptr(k - 1), sift.count(k - 1)
            j.prev_rect[:] = rects[k - 1]
            j.d_pts, j.d_des, j.n_cur = sift.pts_ptr(k), sift.des_ptr(k), sift.count(k)
            j.d_priors_out = pri_ptr
        if k < n_truth:
            j.d_truth, j.truth_stride, j.truth_channels = truth_ptr + k * H * W, W, 1
            j.d_counts = counts_ptr + 16 * k
        a._advance(blend, None, None, quiet=True)

    b = _bare_masker(frame_numbers, multi, index=index)
    fast = fastseq.build_jobs(b, n, rects, frames_ptr, fb, arena, sift, pri_ptr, truth_ptr, n_truth, H, W, counts_ptr)
    assert fast.dtype.itemsize == capi.C.sizeof(capi.FrameJob) and fast.flags.c_contiguous
    assert fast.tobytes() == bytes(jobs)
    assert (b.index, b.current_model) == (a.index, a.current_model)


def test_truth_boxes_are_the_bounding_rectangles_of_the_truth_masks():
    """providers.truth_boxes (OpenCV threshold + boundingRect) against its definition: min / max of the coordinates
    where gray > 127, the previous box when the truth is empty; BGR and single-plane truth frames."""
    from pcm import providers
    rng = np.random.RandomState(11)
    frames = []
    for k in range(24):
        g = np.zeros((60, 90), np.uint8)
        if k % 5 != 3:                                       # every fifth frame: no truth at all
            y, x = rng.randint(0, 50), rng.randint(0, 80)
            g[y:y + rng.randint(1, 10), x:x + rng.randint(1, 10)] = rng.choice([128, 200, 255])
            g[rng.randint(0, 60), rng.randint(0, 90)] = 127  # at the threshold: not foreground
            g[rng.randint(0, 60), rng.randint(0, 90)] = 255  # a lone pixel stretches the box
        frames.append(g)
    want, prev = [], (1, 2, 3, 4)
    for g in frames:
        ys, xs = np.nonzero(g > 127)
        if len(xs):
            prev = (int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1))
        want.append(prev)
    assert providers.truth_boxes(frames, (1, 2, 3, 4)) == want
    bgr = [np.stack([g, 255 - g, g // 2], axis=2) for g in frames]            # channel 0 decides
    assert providers.truth_boxes(bgr, (1, 2, 3, 4)) == want
    assert providers.truth_boxes([np.asfortranarray(g) for g in frames], (1, 2, 3, 4)) == want
