"""Two ranks on two B200s over NCCL: the same checks as test_dist_gloo.py on the product library (skipped on a 1-GPU box)."""
import ctypes as C
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dist_gloo import _free_port  # noqa: E402

pytestmark = pytest.mark.gpu


def _worker_simple(rank, world, port, data, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from bigseqkit_b200 import dist as bd
        from bigseqkit_b200.api import Operator
        cuts = bd.shard_bounds(data, world)
        shard = data[cuts[rank]:cuts[rank + 1]]
        with Operator("Stats", {"Tabular": True, "All": True}, device=rank) as op:
            op.call(shard, partition_id=rank)
            bd.stats_allreduce(op, device=dev)
            row = op.stats_render()
        t = torch.frombuffer(bytearray(shard) + bytearray(64), dtype=torch.uint8).to(dev)
        with Operator("RmDup", {"BySeq": True}, device=rank) as op:
            out, n_rec = bd.rmdup_union(op, t.data_ptr(), len(shard), device=dev)
            host = (C.c_uint8 * max(int(out.n), 1))()
            if out.n:
                cudart = C.CDLL("libcudart.so")
                rc = cudart.cudaMemcpy(host, C.c_void_p(out.data), C.c_size_t(out.n), 2)  # cudaMemcpyDeviceToHost
                assert rc == 0
            kept = bytes(host)[: out.n]
        off, total = bd.output_offsets(len(kept), device=dev)
        ret[rank] = (row, off, total, kept, n_rec)
    finally:
        dist.destroy_process_group()


def test_two_gpus_stats_and_rmdup():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import oracle
    from bigseqkit_b200 import synth
    data = synth.fastq_reads(4 << 20, seed=52, dup_frac=0.2).tobytes()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_simple, args=(2, _free_port(), data, ret), nprocs=2, join=True)
    exp_row = oracle.stats(data, {"Tabular": True, "All": True})[1]
    exp, _, _ = oracle.rmdup(data, {"BySeq": True})
    assert ret[0][0] == exp_row and ret[1][0] == exp_row
    merged = bytearray(ret[0][2])
    for r in range(2):
        _, off, total, kept, _ = ret[r]
        assert total == len(exp)
        merged[off:off + len(kept)] = kept
    assert bytes(merged) == exp
