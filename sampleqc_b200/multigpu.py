"""Multi-GPU plumbing: one process per GPU, cells sharded by sample.

torch.distributed is used ONLY for bootstrap and result assembly (exchange of the 64-byte IPC handles
of the peer-memory windows, gathering alpha_j / p_jk / z on rank 0).  Every per-iteration exchange of the
EM itself happens inside the CUDA kernels over NVLink peer memory (csrc/sqc_xchg.cuh).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A


def shard_by_sample(groups: np.ndarray, J: int, world: int):
    """contiguous sample ranges with balanced cell counts: [(j0, j1, i0, i1)] per rank; groups non-decreasing"""
    groups = np.asarray(groups)
    if groups.size and np.any(np.diff(groups) < 0):
        raise ValueError("multi-GPU sharding needs cells ordered by sample (groups non-decreasing)")
    bounds = np.searchsorted(groups, np.arange(J + 1))
    N = groups.shape[0]
    cuts = [0]
    for r in range(1, world):
        c = int(np.searchsorted(bounds, N * r / world))
        cuts.append(min(max(c, cuts[-1]), J))
    cuts.append(J)
    shards = [(cuts[r], cuts[r + 1], int(bounds[cuts[r]]), int(bounds[cuts[r + 1]])) for r in range(world)]
    # every rank computes the same shards from the same inputs, so all of them fail here together: a rank without
    # samples or cells would fail alone inside the library while its peers wait for its first exchange
    for r, (j0, j1, i0, i1) in enumerate(shards):
        if j1 <= j0 or i1 <= i0:
            raise ValueError(f"rank {r} of {world} would get no cells ({J} samples, {N} cells): use fewer GPUs")
    return shards


def gather_handles(dist, handle: bytes, world: int, device=None) -> bytes:
    """all-gather the ranks' 64-byte IPC handles (works with nccl or gloo)"""
    import torch
    t = torch.tensor(list(handle), dtype=torch.uint8, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return b"".join(bytes(o.cpu().tolist()) for o in out)


def exchange_kwargs(dist, local_device: int, rank: int, world: int, cell_offset: int, N_total: int) -> dict:
    """export this rank's window, gather everybody's handle; returns the Session / make_fit_args keywords"""
    import torch
    from ._lib import check, lib
    buf = C.create_string_buffer(A.SQC_IPC_HANDLE_BYTES)
    err = C.create_string_buffer(512)
    check(lib().sqc_xchg_export(int(local_device), buf, err, 512), err)
    dev = torch.device("cuda", local_device) if dist.get_backend() == "nccl" else None
    handles = gather_handles(dist, buf.raw, world, dev)
    return dict(rank=rank, world=world, cell_offset=int(cell_offset), N_total=int(N_total), peer_handles=handles)


def assemble(dist, rank: int, world: int, fit: dict, shard, J: int, N: int):
    """rank 0 gets the full Rcpp-style list: global parameters from its own copy (identical on all ranks),
    per-sample rows and per-cell labels concatenated in rank order; other ranks get None"""
    per_rank = {"alpha_j": fit["alpha_j"], "p_jk": fit["p_jk"], "z": fit["z"], "n_zero_like": fit["n_zero_like"]}
    if "gamma_i" in fit:
        per_rank["gamma_i"] = fit["gamma_i"]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(per_rank, gathered, dst=0)
    if rank != 0:
        return None
    out = dict(fit)
    out["J"], out["N"] = J, N
    out["alpha_j"] = np.concatenate([g["alpha_j"] for g in gathered], axis=0)
    out["p_jk"] = np.concatenate([g["p_jk"] for g in gathered], axis=0)
    out["z"] = np.concatenate([g["z"] for g in gathered], axis=0)
    out["n_zero_like"] = int(sum(g["n_zero_like"] for g in gathered))
    if "gamma_i" in fit:
        out["gamma_i"] = np.concatenate([g["gamma_i"] for g in gathered], axis=0)
    return out


def fit_robust_sharded(dist, local_device, rank, world, x_local, init_z_local, groups_local, D, J_local, K, N_local,
                       cell_offset, N_total, em_iters, mcd_alpha, mcd_iters, seed, rng_mode=A.SQC_RNG_R_COMPAT, alloc=None, track=False, out=None):
    """this rank's part of a sharded fit_sampleqc_robust_cpp; every rank must call it (returns the local list)"""
    from ._lib import check, lib
    kw = exchange_kwargs(dist, local_device, rank, world, cell_offset, N_total)
    args, keep = A.make_fit_args(x_local, groups_local, D, J_local, K, N_local, em_iters, init_z=init_z_local, mcd_alpha=mcd_alpha,
                                 mcd_iters=mcd_iters, seed=seed, track=track, rng_mode=rng_mode, device=local_device, **kw)
    bufs = out if out is not None else A.FitBuffers(D, J_local, K, N_local, em_iters, A.SQC_VARIANT_ROBUST, track, alloc=alloc)
    err = C.create_string_buffer(512)
    check(lib().sqc_fit_robust(C.byref(args), C.byref(bufs.out), err, 512), err)
    return bufs.as_list()


def fit_mle_sharded(dist, local_device, rank, world, x_local, init_gamma_local, groups_local, D, J_local, K, N_local,
                    cell_offset, N_total, n_iter, seed, alloc=None):
    from ._lib import check, lib
    kw = exchange_kwargs(dist, local_device, rank, world, cell_offset, N_total)
    args, keep = A.make_fit_args(x_local, groups_local, D, J_local, K, N_local, n_iter, init_gamma=init_gamma_local, seed=seed,
                                 device=local_device, **kw)
    bufs = A.FitBuffers(D, J_local, K, N_local, n_iter, A.SQC_VARIANT_MLE, alloc=alloc)
    err = C.create_string_buffer(512)
    check(lib().sqc_fit_mle(C.byref(args), C.byref(bufs.out), err, 512), err)
    return bufs.as_list()
