"""Multi-GPU entry points of the C ABI (dpfhe_multi_*, dpfhe_ipc_*, peer-written outputs) and the call-ordering rule.

Shard equality (SURVEY.md §4 implication iii, §8e): the G-way result equals the one-way result and the oracle's byte for
byte.  On a box with a single GPU the shards are logical (the same device listed twice), as the survey prescribes; with
several GPUs they are real devices and the gather goes through peer stores over NVLink.  The two-process IPC test needs
two GPUs (the exporting process cannot open its own handle).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev(a, device=0):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda(device)


def host(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.fixture(scope="module")
def dp():
    import deeppowers_b200
    return deeppowers_b200


def device_lists():
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]                      # logical shards on one GPU
    if n >= 2:
        lists += [list(range(min(n, 4)))]            # real devices
    if n >= 8:
        lists += [list(range(8))]
    return lists


@pytest.mark.parametrize("log_n,L,batch", [(12, 2, 7), (13, 4, 37)])
def test_multi_host_shards_equal_single_and_oracle(dp, oracle_mod, log_n, L, batch):
    o = oracle_mod.Oracle(log_n, L)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    a = o.fill_uniform(11, 2 * batch).reshape(batch, 2, L, o.N)
    b = o.fill_uniform(12, 2 * batch).reshape(batch, 2, L, o.N)
    want = o.ct_mul_relin(a, b, evk)
    single = dp.Context(log_n, L)
    one = np.zeros_like(a)
    single.ct_mul_relin_host(a, b, evk, one)
    single.close()
    assert np.array_equal(one, want)
    for devices in device_lists():
        m = dp.MultiContext(log_n, L, devices=devices)
        assert m.n == len(devices) and m.devices == devices
        shards = [m.shard(batch, r) for r in range(m.n)]
        assert shards[0][0] == 0 and sum(c for _, c in shards) == batch
        assert all(shards[r][0] + shards[r][1] == shards[r + 1][0] for r in range(m.n - 1))
        got = np.zeros_like(a)
        m.ct_mul_relin_host(a, b, evk, got)
        assert np.array_equal(got, want), "multi-device host result differs (devices %s)" % devices
        g = o.galois_elt(3)
        gk = o.keygen_galois(5, 65537, s, g)
        rot = np.zeros_like(a)
        m.rotate_host(a, g, gk, rot)
        assert np.array_equal(rot, o.rotate(a, g, gk))
        m.close()


@pytest.mark.parametrize("log_n,L,batch", [(12, 3, 10), (13, 4, 41)])
def test_multi_gather_through_peer_stores(dp, oracle_mod, log_n, L, batch):
    """device-resident shards; every device's kernel writes its rows of the gathered result on the root device"""
    o = oracle_mod.Oracle(log_n, L)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    a = o.fill_uniform(21, 2 * batch).reshape(batch, 2, L, o.N)
    b = o.fill_uniform(22, 2 * batch).reshape(batch, 2, L, o.N)
    want = o.ct_mul_relin(a, b, evk)
    for devices in device_lists():
        m = dp.MultiContext(log_n, L, devices=devices)
        for root in sorted({0, m.n - 1}):
            a_sh, b_sh, k_sh = [], [], []
            for r, d in enumerate(devices):
                first, count = m.shard(batch, r)
                a_sh.append(dev(a[first:first + count], d))
                b_sh.append(dev(b[first:first + count], d))
                k_sh.append(dev(evk, d))
            out_root = torch.zeros((batch, 2, L, o.N), dtype=torch.int64, device="cuda:%d" % devices[root])
            for d in set(devices):
                torch.cuda.synchronize(d)
            m.ct_mul_relin_gather(a_sh, b_sh, k_sh, out_root, root, batch)
            assert np.array_equal(host(out_root), want), "gathered result differs (devices %s, root %d)" % (devices, root)
        m.close()


def test_calls_on_different_streams_are_ordered(dp, oracle_mod):
    """All calls of a context share its scratch; a call on another stream must wait for the previous one (dpfhe.h).
    A fused launch on a side stream is followed at once by a host-buffer call (which runs on the context's own
    streams and re-prepares the key companions) and by a launch on a third stream: every result must be right."""
    log_n, L, batch = 13, 4, 300
    c, o = dp.Context(log_n, L), oracle_mod.Oracle(log_n, L)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    g = o.galois_elt(1)
    gk = o.keygen_galois(3, 65537, s, g)
    a = o.fill_uniform(31, 2 * batch).reshape(batch, 2, L, o.N)
    b = o.fill_uniform(32, 2 * batch).reshape(batch, 2, L, o.N)
    da, db, dk, dg = dev(a), dev(b), dev(evk), dev(gk)
    out1, out3 = torch.zeros_like(da), torch.zeros_like(da)
    torch.cuda.synchronize()
    s1, s3 = torch.cuda.Stream(), torch.cuda.Stream()
    h2 = np.zeros((8, 2, L, o.N), dtype=np.uint64)
    c.ct_mul_relin(da, db, dk, out1, batch, stream=s1)
    c.rotate_host(a[:8], g, gk, h2)                      # other streams, other key
    c.rotate(da, g, dg, out3, batch, stream=s3)
    c.synchronize()
    assert np.array_equal(host(out1), o.ct_mul_relin(a, b, evk))
    assert np.array_equal(h2, o.rotate(a[:8], g, gk))
    assert np.array_equal(host(out3), o.rotate(a, g, gk))
    # rotation by slot count: the library derives 5^k mod 2N itself
    out4 = torch.zeros_like(da)
    c.rotate_steps(da, 1, dg, out4, batch)
    torch.cuda.synchronize()
    assert torch.equal(out4, out3)
    c.close()


def test_numa_placed_pinned_memory(dp):
    c = dp.Context(12, 1)
    buf = c.pinned_near(1 << 16)
    assert buf.array.shape == (1 << 16,) and not buf.array.any()
    assert buf.node in (-1, c.numa_node())
    x = np.arange(1 << 12, dtype=np.uint64) % c.moduli[0]
    buf.array[: x.size] = x
    view = buf.array[: x.size]
    c.ntt_fwd_host(view)
    c.ntt_inv_host(view)
    assert np.array_equal(view, x)
    buf.close()
    c.close()


IPC_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import deeppowers_b200 as dp
from oracle import Oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
log_n, L, B = 13, 4, 24
c, o = dp.Context(log_n, L, device=rank), Oracle(log_n, L)
evk = o.keygen_relin(2, 65537, o.keygen_secret(1))
a = o.fill_uniform(41, 2 * B * world).reshape(world * B, 2, L, o.N)
b = o.fill_uniform(42, 2 * B * world).reshape(world * B, 2, L, o.N)
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x).view(np.int64)).cuda()
ct_bytes = 2 * L * o.N * 8
root = c.device_alloc(world * B * ct_bytes) if rank == 0 else None
h = [c.ipc_export(root) if rank == 0 else None]
dist.broadcast_object_list(h, src=0)
base = root if rank == 0 else c.ipc_open(h[0])
da, db, dk = dev(a[rank * B:(rank + 1) * B]), dev(b[rank * B:(rank + 1) * B]), dev(evk)
torch.cuda.synchronize()
c.ct_mul_relin(da, db, dk, base + rank * B * ct_bytes, B)
c.synchronize()
dist.barrier()
if rank == 0:
    class M:
        def __init__(s, p, n): s.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (p, False), "version": 2}
    got = torch.as_tensor(M(root, world * B * ct_bytes // 8), device="cuda").cpu().numpy().view(np.uint64).reshape(a.shape)
    assert np.array_equal(got, o.ct_mul_relin(a, b, evk)), "IPC-gathered result differs from the oracle"
    print("IPC_GATHER_OK")
dist.barrier()
if rank != 0:
    c.ipc_close(base)
dist.barrier()
if rank == 0:
    c.device_free(root)
c.close()
dist.destroy_process_group()
"""


def test_gather_across_processes_through_ipc(tmp_path):
    """one process per GPU (the bench's layout): rank 1's kernel writes its rows into rank 0's exported buffer"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs: a process cannot open its own IPC handle")
    script = tmp_path / "ipc_worker.py"
    script.write_text(IPC_WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "IPC_GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_multi_grouped_host_shards_equal_oracle(dp, oracle_mod):
    """special-prime key switching sharded over the devices (dpfhe_multi_ct_mul_relin_grouped_host): every device list gives the
    oracle's bits, ragged shards included"""
    log_n, L, K, batch, t = 12, 6, 2, 11, 65537
    o = oracle_mod.Oracle(log_n, L)
    Lq = L - K
    oq = oracle_mod.Oracle(log_n, Lq, o.moduli[:Lq])
    a = oq.fill_uniform(3, 2 * batch).reshape(batch, 2, Lq, o.N)
    b = oq.fill_uniform(4, 2 * batch).reshape(batch, 2, Lq, o.N)
    key = o.fill_uniform(5, 2 * o.grouped_digits(K)).reshape(-1, 2, L, o.N)
    exp = o.ct_mul_relin_grouped(K, a, b, key, t)
    for devices in device_lists():
        m = dp.MultiContext(log_n, L, devices=devices)
        out = np.zeros_like(a)
        m.ct_mul_relin_grouped_host(K, a, b, key, out, t)
        assert np.array_equal(out, exp), devices
        m.close()
