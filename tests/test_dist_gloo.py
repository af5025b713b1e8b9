"""CPU tier: the N>1 plumbing (frame/sequence sharding + pose gather + serial chain replay) with world_size-2 gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from openvo_b200 import dist as D
    try:
        # sequences sharded round-robin; each rank fabricates per-frame relative transforms for its sequences
        n_seq, n_frames = 5, 6
        mine = D.shard_sequences(n_seq, rank, world)
        rng = np.random.default_rng(7)
        allT = rng.normal(0, 0.01, (n_seq, n_frames, 4, 4)) + np.eye(4)
        allT[..., 3, :] = [0, 0, 0, 1]
        status = np.ones((n_seq, n_frames), np.int32)
        status[1, 3] = 0  # a failed frame is not committed
        local_T = allT[mine]
        local_s = status[mine]
        T, s, owner = D.gather_poses(local_T, local_s, mine, n_seq)
        assert np.array_equal(T, allT) and np.array_equal(s, status)
        status[2, 4] = 2  # a fall-back alignment replaces the previous frame's transform
        chains = D.replay_chains(T, status)
        ref = []
        for i in range(n_seq):
            c, p = np.eye(4), np.eye(4)
            for j in range(n_frames):
                if status[i, j] == 1:
                    p, c = c, allT[i, j] @ c
                elif status[i, j] == 2:
                    p, c = c, allT[i, j] @ p
            ref.append(c)
        assert np.abs(chains - np.stack(ref)).max() < 1e-12
        # ragged: rank 1 holds two frames fewer than rank 0 (its missing frames read as "not committed")
        nf = n_frames - 2 * rank
        T2, s2, _ = D.gather_poses(allT[mine][:, :nf], status[mine][:, :nf], mine, n_seq)
        assert T2.shape[1] == n_frames
        for i in range(n_seq):
            k = n_frames - 2 * (i % world)
            assert np.array_equal(T2[i, :k], allT[i, :k]) and np.array_equal(s2[i, :k], status[i, :k]) and not s2[i, k:].any()
        # frame-chunk gather: every rank contributes its contiguous chunk of one sequence
        nfr = 7
        first, start, end = D.shard_frames(nfr, rank, world)
        Tc = allT[0, :1].repeat(end - start, 0) * (np.arange(start, end)[:, None, None] + 1)
        sc = np.ones(end - start, np.int32)
        Tg, sg = D.gather_frame_chunks(start, Tc, sc, nfr)
        assert sg.tolist() == [1] * nfr and np.allclose(Tg[5], allT[0, 0] * 6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gather_and_replay_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_helpers():
    from openvo_b200 import dist as D
    assert D.shard_sequences(5, 0, 2) == [0, 2, 4] and D.shard_sequences(5, 1, 2) == [1, 3]
    chunks = [D.shard_frames(10, r, 3) for r in range(3)]
    # contiguous chunks with a one-frame halo on the left
    assert chunks[0] == (0, 0, 4) and chunks[1] == (3, 4, 8) and chunks[2] == (7, 8, 10)
