"""BASELINE configs[3] on N GPUs: channelize (2^16-point STFT) -> detect -> bin 64 fine channels
-> fold into 1024 phase bins x 1024 channels; time slices sharded over ranks, ONE all-reduce of
the folded profile and counts over NCCL.  Run with torchrun; prints one line of timings (rank 0).

    python -m torch.distributed.run --nproc-per-node N scripts/cfg4_pipeline.py [log2_samples_per_rank [npol [raw]]]

``raw`` = int8 | u4 | u2 feeds the channelizer with raw baseband (decode fused into its first pass)
and adds an end-to-end line: every step uploads its block from pinned host memory and reads the
folded profile back.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pulsarbat_b200 as pb  # noqa: E402
from pulsarbat_b200 import sharding  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ["PBK_DEVICE"] = str(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
npol = int(sys.argv[2]) if len(sys.argv) > 2 else 2
raw = sys.argv[3] if len(sys.argv) > 3 else None
n_per_rank, nper, fsum, nbin = 2 ** lg, 2 ** 16, 64, 1024
sr = 400e6
g = torch.Generator(device=dev)
g.manual_seed(16 + rank)
raw_kw = {}
if raw is None:
    x = torch.randn((n_per_rank, 1, npol, 2), device=dev, dtype=torch.float32, generator=g)
    xd = pb.DeviceArray(torch.view_as_complex(x) if npol > 1 else torch.view_as_complex(x)[:, :, 0])
else:
    rshape = {"int8": (n_per_rank, 1, npol, 2), "u4": (n_per_rank, 1, npol),
              "u2": (n_per_rank, npol // 2)}[raw]
    if raw == "int8":
        x = torch.randint(-127, 128, rshape, device=dev, dtype=torch.int8, generator=g)
    else:
        x = torch.randint(0, 256, rshape, device=dev, dtype=torch.uint8, generator=g)
    xd = pb.DeviceArray(x)
    raw_kw = {"raw": raw, **({"raw_shape": (1, npol)} if raw == "u2" else {})}
coeffs = [0.123, 29.7, 1e-6]
seg_per_rank = n_per_rank // nper


def step(src=None):
    if raw is None:
        # ONE library call: bins + counts, first FFT pass, last FFT pass whose epilogue detects,
        # sums 64 fine channels and adds the result to the profile row of the segment's phase bin
        prof, cnt = pb.kernels.stft_fold(xd if src is None else src, nper, coeffs, sr / nper, nbin,
                                         freq_sum=fsum, n0=rank * seg_per_rank)
    else:
        zc = pb.kernels.stft(xd if src is None else src, nper, **raw_kw)   # (segments, 65536)
        inten = pb.kernels.detect(zc, freq_sum=fsum)                     # (segments, 1024)
        prof, cnt = pb.kernels.fold(inten, coeffs, sr / nper, nbin, n0=rank * seg_per_rank)
    return sharding.allreduce_profiles(prof, cnt)


for _ in range(3):
    prof, cnt = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    prof, cnt = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
total = int(np.asarray(cnt).sum())
if rank == 0:
    assert total == world * seg_per_rank, (total, world * seg_per_rank)
    print(f"cfg4: {world} GPU(s) x 2^{lg} samples x {npol} pol: {float(t.item()):.3f} ms per step -> "
          f"{world * n_per_rank * npol / float(t.item()) / 1e6:.1f} Gsamples/s; profile "
          f"{tuple(np.asarray(prof).shape)}, counts sum {total} (exact)", flush=True)
if raw is not None:
    # end to end: pinned raw block -> device -> channelize/detect/fold -> profile back on the host
    hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x)
    dbuf = [torch.empty_like(x) for _ in range(2)]
    s_up = torch.cuda.Stream(dev)
    up_done = [torch.cuda.Event() for _ in range(2)]
    used = [torch.cuda.Event() for _ in range(2)]

    def e2e(nblk):
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(s_up):
            dbuf[0].copy_(hx, non_blocking=True)
            up_done[0].record(s_up)
        acc = 0.0
        for i in range(nblk):
            sl = i & 1
            if i + 1 < nblk:
                with torch.cuda.stream(s_up):
                    if i >= 1:
                        s_up.wait_event(used[sl ^ 1])
                    dbuf[sl ^ 1].copy_(hx, non_blocking=True)
                    up_done[sl ^ 1].record(s_up)
            cur.wait_event(up_done[sl])
            p, c = step(pb.DeviceArray(dbuf[sl]))
            used[sl].record(cur)
            acc += float(np.asarray(p).ravel()[0])        # the profile of every block reaches the host
        return acc

    e2e(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e(K)
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / K * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"cfg4 e2e ({raw} blocks of {hx.numel() * hx.element_size() / 2**20:.0f} MiB from pinned "
              f"host memory, upload overlapped): {float(dt.item()):.3f} ms per step -> "
              f"{world * n_per_rank * npol / float(dt.item()) / 1e6:.1f} Gsamples/s", flush=True)
if world > 1:
    dist.destroy_process_group()
