"""Two ranks on two B200s: the exchange steps through the C ABI over NCCL (bsk_comm_init, bsk_stats_allreduce,
bsk_rmdup_sharded, bsk_output_offsets).  Needs 2 GPUs (`gpurun --gpus 2`); NCCL refuses two ranks on one device, so on
a 1-GPU box the same exchange logic is covered by tests/test_exchange.py (single-process form) instead."""
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


def _worker(rank, world, port, data, ret):
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
            bd.init_comm(op)
            assert op.comm_rank() == (rank, world)
            op.call(shard, partition_id=rank)
            bd.stats_allreduce(op, device=dev)
            row = op.stats_render()
        t = torch.frombuffer(bytearray(shard) + bytearray(64), dtype=torch.uint8).to(dev)
        with Operator("RmDup", {"BySeq": True}, device=rank) as op:
            bd.init_comm(op)
            out, n_rec = bd.rmdup_union(op, t.data_ptr(), len(shard), device=dev)
            kept = op.fetch(out)[0].tobytes()
            off, total = bd.output_offsets(len(kept), device=dev, op=op)
        ret[rank] = (row, off, total, kept, n_rec)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mib", [4, 192])
def test_two_gpus_stats_and_rmdup(mib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); tests/test_exchange.py covers the exchange logic on one")
    import oracle
    from bigseqkit_b200 import synth
    arr, _ = synth.native_fastq(mib << 20, seed=52, dup_frac=0.2)
    data = arr.tobytes()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), data, ret), nprocs=2, join=True)
    sopts = {"Tabular": True, "All": True}
    exp_row = oracle.run_mt_full("stats", arr.ctypes.data, arr.nbytes, sopts, os.cpu_count() or 1)["row"]
    exp = oracle.run_mt_full("rmdup", arr.ctypes.data, arr.nbytes, {"BySeq": True}, os.cpu_count() or 1)["data"].tobytes()
    assert ret[0][0] == exp_row and ret[1][0] == exp_row
    merged = bytearray(ret[0][2])
    for r in range(2):
        _, off, total, kept, _ = ret[r]
        assert total == len(exp)
        merged[off:off + len(kept)] = kept
    assert bytes(merged) == exp
