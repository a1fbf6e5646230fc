"""Host-side logic of the data-parallel path at world size 2 on CPU with the gloo backend:
weight replication from rank 0, the flat gradient buffer every .grad is a view of, the single averaging
all-reduce, identical patch ids on every rank, and the global-batch normalisation of the masked L1 terms.
No kernel runs here (the compute path is CUDA-only); this covers what surrounds it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from dfmir_b200 import registration_model as rm
        torch.manual_seed(100 + rank)                    # different initial weights per rank
        opt = rm.default_options(batch_size=2, ngf=8, crop_size=32, load_size=32, netF_nc=16, num_patches=8, gpu_ids=[])
        m = rm.REGISTRATIONModel(opt)
        m.netF.create_mlp([torch.zeros(1, c, 2, 2) for c in (1, 16, 32, 32, 32)])
        before = torch.cat([p.detach().flatten() for n in m.model_names for p in getattr(m, 'net' + n).parameters()])
        m.parallelize()
        after = torch.cat([p.detach().flatten() for n in m.model_names for p in getattr(m, 'net' + n).parameters()])
        gathered = [torch.empty_like(after) for _ in range(world)]
        dist.all_gather(gathered, after)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "weights differ across ranks after parallelize()"
        if rank == 0:
            assert torch.equal(before, after), "rank 0 is the source of the broadcast"
        else:
            assert not torch.equal(before, after)

        # every .grad is a view of one flat buffer, laid out bucket by bucket in the order the backward pass
        # completes them ([R, F] | G decoder | G encoder); the buckets' all-reduces average all of them
        all_params = [p for n in m.model_names for p in getattr(m, 'net' + n).parameters() if p.requires_grad]
        groups = m._grad_bucket_groups()
        params = [p for g in groups for p in g]
        assert len(groups) == 3 and sorted(map(id, params)) == sorted(map(id, all_params))
        enc_ids = {id(p) for i, mod in enumerate(m.netG.model) if i <= 16 for p in mod.parameters()}
        assert {id(p) for p in groups[2]} == enc_ids, "the last bucket holds the generator layers the encoder passes share"
        assert m._flat_grad is not None and m._flat_grad.numel() == sum(p.numel() for p in params)
        base = m._flat_grad.data_ptr()
        off = 0
        for p in params:
            assert p.grad.data_ptr() == base + 4 * off and p.grad.shape == p.shape
            off += p.numel()
        assert [b[0].numel() for b in m._buckets] == [sum(p.numel() for p in g) for g in groups]
        for i, p in enumerate(params):
            p.grad.fill_(float(rank + 1) * (i + 1))
        m._sync_grads()
        for i, p in enumerate(params):
            assert torch.allclose(p.grad, torch.full_like(p.grad, (1 + world) / 2.0 * (i + 1)))
        m._zero_grads()
        assert float(m._flat_grad.abs().sum()) == 0.0 and all(float(p.grad.abs().sum()) == 0.0 for p in params)

        # the overlapped path: post-accumulate hooks launch each bucket's all-reduce when its last gradient lands
        m._zero_grads()
        loss = sum((p * float(rank + 1)).sum() for p in params)
        loss.backward()
        assert len(m._bucket_work) == len(groups), "every bucket's collective is issued during backward"
        m._sync_grads()
        assert not m._bucket_work
        for p in params:
            assert torch.allclose(p.grad, torch.full_like(p.grad, (1 + world) / 2.0))
        m._zero_grads()

        # identical patch ids on every rank (one permutation per layer shared by the whole batch, networks.py:609)
        # from PatchSampleF's dedicated generator; the global RNG (data order, augmentation) stays per-rank
        ids = torch.randperm(1000, generator=m.netF.generator)[:16]
        got = [torch.empty_like(ids) for _ in range(world)]
        dist.all_gather(got, ids)
        assert all(torch.equal(g, got[0]) for g in got), "patch ids differ across ranks"
        glob = torch.randperm(1000)[:16]
        got = [torch.empty_like(glob) for _ in range(world)]
        dist.all_gather(got, glob)
        assert not torch.equal(got[0], got[1]), "parallelize() must not reseed the global generator"

        # masked L1: rank-local means rescaled so that the rank average is the global-batch masked mean
        S = torch.tensor([3.0, 10.0])[rank]              # sum |a-b| * mask on this rank
        M = torch.tensor([4.0, 16.0])[rank]              # mask count on this rank
        local = S / M * rm.global_mask_scale(M.clone(), world)
        dist.all_reduce(local)
        assert abs(float(local) / world - 13.0 / 20.0) < 1e-6
        # a rank with an empty mask contributes nothing and does not produce NaN
        M0 = torch.tensor([0.0, 5.0])[rank]
        sc = rm.global_mask_scale(M0.clone(), world)
        assert torch.isfinite(sc) and (float(sc) == 0.0 if rank == 0 else abs(float(sc) - 2.0) < 1e-6)
        # both masked terms of the step through one collective
        both = rm.global_mask_scale(torch.stack([M, M0]), world)
        assert abs(float(both[0]) - float(M) * world / 20.0) < 1e-6 and abs(float(both[1]) - float(sc)) < 1e-6
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.timeout(300)
def test_data_parallel_host_logic_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, res in results:
        assert res == "ok", f"rank {rank}:\n{res}"
