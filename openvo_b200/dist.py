"""Multi-GPU plumbing: one process per GPU, work sharded by sequence (or by contiguous frame chunk with a
one-frame halo), no collective on the data path; the only exchange is one all-gather of the per-frame relative transforms
(4x4 f64 = 128 B) and status words at the end of a chunk, after which the ordered pose chain is replayed serially
(ref: src/openVO/stereo_odometer.py:136-160 — only the pose chain is ordered).  NCCL over NVLink on GPUs, gloo in the CPU
tests; the message is a few tens of KB, so this is latency-only and done once per chunk, never per frame.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_sequences(n_seq, rank, world):
    """Round-robin assignment of independent sequences to ranks."""
    return list(range(rank, n_seq, world))


def shard_frames(n_frames, rank, world):
    """Contiguous chunk of one sequence for this rank -> (first frame to process incl. the one-frame halo, first frame owned,
    end).  The halo frame's features are recomputed rather than shipped (64 KB of descriptors + the disparity map)."""
    per = (n_frames + world - 1) // world
    start, end = min(rank * per, n_frames), min((rank + 1) * per, n_frames)
    return (max(start - 1, 0), start, end)


def _device():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_poses(local_T, local_status, mine, n_seq):
    """All-gather per-frame relative transforms.

    local_T: float64 [len(mine), n_frames, 4, 4]; local_status: int32 [len(mine), n_frames] (1 = committed);
    mine: global sequence ids held by this rank.  Returns (T [n_seq, n_frames, 4, 4], status [n_seq, n_frames],
    owner [n_seq]) on every rank.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    n_local = n_frames = local_T.shape[1] if len(mine) else 0
    dev = _device()
    meta = torch.tensor([len(mine), n_frames], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    cap = int(max(int(m[0]) for m in metas))
    n_frames = int(max(int(m[1]) for m in metas))
    # one packed message per rank: ids | status | transforms
    pack = torch.zeros(cap * (1 + n_frames + n_frames * 16), dtype=torch.float64, device=dev)
    if len(mine):
        k = len(mine)
        buf = np.zeros((cap, 1 + n_frames + n_frames * 16))
        buf[:, 0] = -1
        buf[:k, 0] = mine
        # a rank may hold fewer frames than the longest one: pad with status 0 (frame not committed) / zero transforms
        buf[:k, 1:1 + n_local] = local_status
        padded = np.zeros((k, n_frames, 16))
        padded[:, :n_local] = np.asarray(local_T, np.float64).reshape(k, n_local, 16)
        buf[:k, 1 + n_frames:] = padded.reshape(k, -1)
        pack = torch.from_numpy(buf.reshape(-1)).to(dev)
    else:
        pack.view(cap, -1)[:, 0] = -1
    packs = [torch.zeros_like(pack) for _ in range(world)]
    dist.all_gather(packs, pack)
    T = np.zeros((n_seq, n_frames, 4, 4))
    status = np.zeros((n_seq, n_frames), np.int32)
    owner = np.full(n_seq, -1, np.int32)
    for r, p in enumerate(packs):
        rows = p.cpu().numpy().reshape(cap, -1)
        for row in rows:
            sid = int(row[0])
            if sid < 0:
                continue
            owner[sid] = r
            status[sid] = row[1:1 + n_frames].astype(np.int32)
            T[sid] = row[1 + n_frames:].reshape(n_frames, 4, 4)
    return T, status, owner


def replay_chains(T, status):
    """Serial replay of the pose chain of every sequence (ref: stereo_odometer.py:136-150).  status: 0 = frame not committed,
    1 = T aligns the frame to the last committed one (c_T_w <- T @ c_T_w), 2 = fall-back alignment to the frame before it
    (c_T_w <- T @ c_T_w_prev, the reference's second attempt)."""
    out = np.tile(np.eye(4), (T.shape[0], 1, 1))
    for s in range(T.shape[0]):
        cur, prev = np.eye(4), np.eye(4)
        for t in range(T.shape[1]):
            if status[s, t] == 1:
                prev, cur = cur, T[s, t] @ cur
            elif status[s, t] == 2:
                prev, cur = cur, T[s, t] @ prev
        out[s] = cur
    return out


def run_frame_chunk(odometer, lefts, rights, rank, world):
    """BASELINE config 4: one sequence sharded by contiguous frame chunk.  This rank runs `odometer` (a fresh StereoOdometer)
    over its chunk preceded by a one-frame halo and returns (first owned frame, T [n,4,4], status [n]) for the frames it owns.
    Exact w.r.t. the sequential reference whenever no frame inside the first two frames of a chunk is skipped (the skip state
    machine B4 looks two committed frames back); `status` lets the caller detect that case and re-run those frames."""
    n_frames = len(lefts)
    first, start, end = shard_frames(n_frames, rank, world)
    T = np.tile(np.eye(4), (end - start, 1, 1))
    status = np.zeros(end - start, np.int32)
    for f in range(first, end):
        ok = odometer.update(lefts[f], rights[f])
        if f >= start and f > 0:
            status[f - start] = odometer.last_mode if ok else 0
            if ok and odometer.last_T is not None and odometer.last_mode:
                T[f - start] = odometer.last_T
    return start, T, status


def gather_frame_chunks(start, T, status, n_frames):
    """All-gather the per-chunk results of run_frame_chunk -> (T [n_frames,4,4], status [n_frames]) on every rank."""
    world = dist.get_world_size()
    dev = _device()
    per = (n_frames + world - 1) // world
    pack = torch.zeros(per * 18, dtype=torch.float64, device=dev)
    buf = np.zeros((per, 18))
    buf[:, 0] = -1
    k = len(status)
    buf[:k, 0] = np.arange(start, start + k)
    buf[:k, 1] = status
    buf[:k, 2:] = T.reshape(k, 16)
    pack.copy_(torch.from_numpy(buf.reshape(-1)))
    packs = [torch.zeros_like(pack) for _ in range(world)]
    dist.all_gather(packs, pack)
    Tall = np.tile(np.eye(4), (n_frames, 1, 1))
    sall = np.zeros(n_frames, np.int32)
    for p in packs:
        for row in p.cpu().numpy().reshape(per, 18):
            if row[0] >= 0:
                f = int(row[0])
                sall[f] = int(row[1])
                Tall[f] = row[2:].reshape(4, 4)
    return Tall, sall
