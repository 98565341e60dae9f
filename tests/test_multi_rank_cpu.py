"""N > 1 host logic on the CPU: two gloo ranks shard a two-round render (round sharding) and a one-round render
(tile sharding), sum their partial framebuffers with the package's reduce, and rank 0 checks the result against a
single-process render with the same seeds.  The renderer inside the ranks is the CPU oracle (test infrastructure) --
what is under test is the sharding arithmetic, the seed bases and the reduce."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import checkers
    from rgk_b200 import abi, multigpu, scenes
    O = checkers.oracle()
    pack, cfg = scenes.load_builtin("cornell-box", width=64, height=64, multisample=2, recursion_max=4)
    desc = pack.desc()
    h = O.scene_create(desc)
    ca = cfg.camera_args()
    cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
    p = cfg.params()
    tasks = O.generate_tasks(32, 64, 64)
    nt = len(tasks)
    # (a) round sharding: step 0 -> rank r renders round r
    rnd = multigpu.round_for_rank(0, rank, world)
    fb, cnt, _ = O.render_round(h, cam, p, tasks, seedstart=42, seedcount_base=multigpu.seedcount_base(rnd, nt), nthreads=2)
    tfb, tcnt = torch.from_numpy(fb), torch.from_numpy(cnt.astype(np.int32))
    multigpu.reduce_framebuffer(tfb, tcnt)
    # (b) tile sharding of round 0: tile i keeps seed index i
    fb2 = np.zeros_like(fb); cnt2 = np.zeros_like(cnt)
    for i in multigpu.tiles_for_rank(nt, rank, world):
        one = (abi.Task * 1)(tasks[i])
        O.render_round(h, cam, p, one, seedstart=42, seedcount_base=i, fb=(fb2, cnt2), nthreads=1)
    tfb2, tcnt2 = torch.from_numpy(fb2), torch.from_numpy(cnt2.astype(np.int32))
    multigpu.reduce_framebuffer(tfb2, tcnt2)
    if rank == 0:
        ref, rc, _ = O.render_round(h, cam, p, tasks, seedstart=42, seedcount_base=0, nthreads=2)
        one_round = ref.copy()
        ref, rc, _ = O.render_round(h, cam, p, tasks, seedstart=42, seedcount_base=nt, fb=(ref, rc), nthreads=2)
        ok_a = np.array_equal(tfb.numpy(), ref) and np.array_equal(tcnt.numpy().astype(np.uint32), rc) and int(rc.min()) == 2 * p.multisample
        ok_b = np.array_equal(tfb2.numpy(), one_round) and int(tcnt2.min()) == p.multisample
        np.save(out, np.array([ok_a, ok_b]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_round_and_tile_sharding(tmp_path):
    out = str(tmp_path / "ok.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ok = np.load(out)
    assert ok[0], "round-sharded sum differs from the two-round single-process render"
    assert ok[1], "tile-sharded sum differs from the one-round single-process render"


def test_shard_arithmetic():
    from rgk_b200 import multigpu
    assert [multigpu.round_for_rank(s, r, 4) for s in range(2) for r in range(4)] == list(range(8))
    parts = [multigpu.tiles_for_rank(10, r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == list(range(10)) and max(len(x) for x in parts) - min(len(x) for x in parts) <= 1
    assert multigpu.seedcount_base(3, 2040) == 6120
