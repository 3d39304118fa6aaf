"""Size-independent properties at the block size of BASELINE.json's configs (1 GiB of 150 bp FASTQ = 3.1 M reads), on the
GPU through the C ABI with device-resident data -- sizes the oracle does not finish in seconds:
  * seq --reverse --complement is an involution (applied twice it restores every byte, element offsets included);
  * stats is invariant under it (same length histogram, Q20 / Q30, gap count), and num_seqs / sum_len match the generator;
  * rmdup --by-seq is idempotent, keeps exactly one record per distinct sequence, and keeps a prefix-closed subset."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_BYTES = 1 << 30


@pytest.fixture(scope="module")
def block():
    import torch
    from bigseqkit_b200 import synth
    host = synth.fastq_reads(N_BYTES, seed=2, dup_frac=0.2)
    dev = torch.device("cuda", 0)
    d = torch.empty(host.nbytes + 64, dtype=torch.uint8, device=dev)
    d[: host.nbytes].copy_(torch.from_numpy(host))
    torch.cuda.synchronize()
    return host, d


def _as_tensor(ptr, n, torch):
    """copy n device bytes at ptr into a fresh torch tensor (the ctx owns the source until its next call)"""
    out = torch.empty(n + 64, dtype=torch.uint8, device="cuda:0")
    rc = C.CDLL("libcudart.so").cudaMemcpy(C.c_void_p(out.data_ptr()), C.c_void_p(ptr), C.c_size_t(n), 3)
    assert rc == 0
    torch.cuda.synchronize()  # a device-to-device cudaMemcpy does not block the host, and ctx streams are non-blocking
    return out


def test_revcomp_is_an_involution_and_stats_invariant(block):
    import torch
    from bigseqkit_b200 import Operator
    host, d = block
    n = host.nbytes
    opts = {"Reverse": True, "Complement": True}
    with Operator("SeqTransform", opts, device=0) as op:
        o1 = op.call_device(d.data_ptr(), n)
        assert o1.n == n and op.timings()["fused_blocks"] == 1
        n_rec = int(o1.n_records)
        t1 = _as_tensor(o1.data, n, torch)
        e1 = _as_tensor(o1.elem_off, 8 * (n_rec + 1), torch)
        o2 = op.call_device(t1.data_ptr(), n)
        assert o2.n == n and int(o2.n_records) == n_rec
        t2 = _as_tensor(o2.data, n, torch)
        e2 = _as_tensor(o2.elem_off, 8 * (n_rec + 1), torch)
    assert not torch.equal(t1[:n], d[:n])
    assert torch.equal(t2[:n], d[:n])
    assert torch.equal(e1[: 8 * (n_rec + 1)], e2[: 8 * (n_rec + 1)])
    # element offsets are the record starts: every one points at an '@' that follows a newline
    offs = e1[: 8 * n_rec].view(torch.int64)
    assert int(offs[0]) == 0 and bool((t1[offs] == ord("@")).all()) and bool((t1[offs[1:] - 1] == 10).all())
    assert int(e1[8 * n_rec: 8 * (n_rec + 1)].view(torch.int64)[0]) == n
    with Operator("Stats", {"All": True}, device=0) as st:
        st.call_device(d.data_ptr(), n)
        a = st.stats_result()
        st.reset()
        st.call_device(t1.data_ptr(), n)
        b = st.stats_result()
    assert a == b
    assert a["num"] == n_rec and a["sum_len"] == 150 * n_rec and a["min_len"] == 150 and a["hist"] == [(150, n_rec)]
    assert a["type"] == "DNA" and 0 < a["q30"] < a["q20"] < a["sum_len"]


def test_rmdup_idempotent_and_counts_distinct_sequences(block):
    import torch
    from bigseqkit_b200 import Operator
    host, d = block
    n = host.nbytes // 2
    k = host[:n].tobytes().rfind(b"\n@SIM:")
    n = k + 1
    # distinct sequences of the prefix, computed on the host from the fixed record layout (line 2 of every record)
    data = host[:n]
    nl = np.flatnonzero(data == 10)
    assert nl.size % 4 == 0
    seq_start = nl[0::4] + 1
    seqs = np.lib.stride_tricks.as_strided(data, shape=(seq_start.size, 150), strides=(0, 1))  # placeholder, replaced below
    idx = seq_start[:, None] + np.arange(150)[None, :]
    seqs = data[idx]
    distinct = np.unique(seqs, axis=0).shape[0]
    with Operator("RmDup", {"BySeq": True}, device=0) as op:
        o1 = op.call_device(d.data_ptr(), n)
        kept1, n1 = int(o1.n_elem), int(o1.n)
        assert int(o1.n_records) == seq_start.size and kept1 == distinct
        assert op.rmdup_removed() == seq_start.size - distinct
        t1 = _as_tensor(o1.data, n1, torch)
        op.reset()
        o2 = op.call_device(t1.data_ptr(), n1)
        assert int(o2.n) == n1 and int(o2.n_elem) == kept1 and op.rmdup_removed() == 0
        t2 = _as_tensor(o2.data, n1, torch)
    assert torch.equal(t1[:n1], t2[:n1])
    # the first record always survives and the output starts like the input
    first_len = int(nl[3]) + 1
    assert torch.equal(t1[:first_len], d[:first_len])
