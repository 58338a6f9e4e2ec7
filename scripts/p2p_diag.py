"""What does peer memory cost on this box?  (torchrun, 2+ ranks)

  * topology as nvidia-smi sees it (NVLink or PCIe between the GPUs?)
  * NCCL all_to_all_single bandwidth (what NCCL's transport reaches)
  * bulk copy out of a CUDA-IPC-mapped peer arena (cudaMemcpy peer -> local)
  * kernel reads of the mapped peer arena: contiguous stream and random 64-byte rows (= the sharded
    gather's access pattern), via torch ops on a tensor aliasing the peer pointer
"""
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from deepctr_torch_b200 import sharded


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda:%d" % local)
    dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        for cmd in (["nvidia-smi", "topo", "-m"], ["nvidia-smi", "nvlink", "--status", "-i", "0"]):
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
                print("$ %s\n%s" % (" ".join(cmd), "\n".join(out.splitlines()[:24])), flush=True)
            except Exception as ex:            # noqa: BLE001
                print("%s failed: %s" % (cmd, ex))
        print("can_device_access_peer(0,1) =", torch.cuda.can_device_access_peer(0, 1), flush=True)
    n = 256 << 20
    # NCCL all-to-all
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    dst = torch.empty_like(src)
    t = timed(lambda: dist.all_to_all_single(dst, src))
    if rank == 0:
        print("NCCL all_to_all_single %d MB per rank: %.2f ms -> %.1f GB/s sent per rank to peers" % (
            n >> 20, t * 1e3, n * (world - 1) / world / t / 1e9), flush=True)
    # peer arena
    arena = sharded.P2PArena(n, dev)
    dist.barrier()
    peer = (rank + 1) % world
    raw = sharded._RawCuda(arena.peer_ptr[peer], n)
    remote = torch.as_tensor(raw, device=dev)
    localbuf = torch.empty(n, dtype=torch.uint8, device=dev)
    t = timed(lambda: localbuf.copy_(remote))
    print("rank %d: bulk copy from the mapped peer arena: %.2f ms -> %.1f GB/s" % (rank, t * 1e3, n / t / 1e9), flush=True)
    rf = remote.view(torch.float32).view(-1, 16)                 # 64-byte rows
    t = timed(lambda: rf.sum())
    print("rank %d: kernel streaming read of the peer arena (reduce): %.2f ms -> %.1f GB/s" % (rank, t * 1e3, n / t / 1e9), flush=True)
    idx = torch.randint(0, rf.shape[0], (65536 * 13,), device=dev)
    out = torch.empty(idx.numel(), 16, device=dev)
    t = timed(lambda: torch.index_select(rf, 0, idx, out=out))
    print("rank %d: random 64-byte rows from the peer arena (index_select, %d rows): %.2f ms -> %.1f GB/s" % (
        rank, idx.numel(), t * 1e3, idx.numel() * 64 / t / 1e9), flush=True)
    lf = localbuf.view(torch.float32).view(-1, 16)
    t = timed(lambda: torch.index_select(lf, 0, idx, out=out))
    print("rank %d: same gather from LOCAL memory: %.2f ms -> %.1f GB/s" % (rank, t * 1e3, idx.numel() * 64 / t / 1e9), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
