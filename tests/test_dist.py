"""world_size-2 gloo test of the multi-GPU host logic (sharding + record gather).  The per-shard compute is the
CPU oracle here (tests may use it as a stand-in); on the GPU box the same code path runs gsmcal.calibrate_batch."""
import math
import os
import sys

import torch.distributed as dist
import torch.multiprocessing as mp


def _oracle_compute(raw, carrier, tpl, coef):
    import gsmcal_oracle as oracle
    from gsmcal._lib import StreamResult
    out = (StreamResult * len(raw))()
    for i, r in enumerate(raw):
        res = oracle.calibrate_stream(r, carrier, tpl, coef)
        out[i].n_coarse = -1 if res["coarse_pos"][0] == -1 else len(res["coarse_pos"])
        out[i].n_fcch = -1 if res["fcch_pos"][0] == -1 else len(res["fcch_pos"])
        out[i].n_pos_info = -1 if res["pos_info"].shape == (1, 2) else len(res["pos_info"])
        out[i].sampling_ppm[0], out[i].sampling_ppm[1] = res["sampling_ppm"]
        out[i].carrier_ppm[0], out[i].carrier_ppm[1] = res["carrier_ppm"]
        out[i].total_sampling_ppm = res["total_sampling_ppm"]
        out[i].total_carrier_ppm = res["total_carrier_ppm"]
    return out


def _worker(rank, world, port, raw, q):
    for p in sys.path_extra:
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gsmcal_oracle as oracle
    from gsmcal import dist as gdist
    tpl = oracle.gsm_SCH_training_sequence_gen(8)
    coef = oracle.fir1(46, 200e3 / (oracle.SYMBOL_RATE * 8))
    recs = gdist.calibrate_sharded(raw, 957.4e6, tpl, coef, _oracle_compute)
    q.put((rank, [(r.n_coarse, r.n_fcch, r.n_pos_info, r.total_sampling_ppm, r.total_carrier_ppm) for r in recs]))
    dist.destroy_process_group()


def test_shard_ranges_cover_all_streams():
    from gsmcal.dist import shard_range
    for n, w in ((1024, 8), (5, 2), (3, 4), (126, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_gloo_gather_matches_single_process():
    from gsmcal import synth
    import gsmcal_oracle as oracle
    n = 400000                                          # short captures: 3-4 FCCH -> sentinel records, plus one noise stream
    specs = [synth.random_spec(s, n) for s in (1, 2)] + [synth.StreamSpec(seed=9, n_samples=n, noise_only=True)]
    raw = synth.generate_batch(specs).numpy()
    tpl = oracle.gsm_SCH_training_sequence_gen(8)
    coef = oracle.fir1(46, 200e3 / (oracle.SYMBOL_RATE * 8))
    single = _oracle_compute(raw, 957.4e6, tpl, coef)
    expect = [(r.n_coarse, r.n_fcch, r.n_pos_info, r.total_sampling_ppm, r.total_carrier_ppm) for r in single]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    sys.path_extra = list(sys.path)
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_entry, args=(r, 2, port, raw, q, list(sys.path))) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        assert len(got[rank]) == 3
        for a, b in zip(got[rank], expect):
            assert a[:3] == b[:3]
            for x, y in zip(a[3:], b[3:]):
                assert (math.isinf(x) and math.isinf(y)) or x == y


def _worker_entry(rank, world, port, raw, q, paths):
    sys.path_extra = paths
    _worker(rank, world, port, raw, q)
