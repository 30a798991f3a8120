"""world_size-2 gloo (CPU) test of the multi-GPU exchange plan (collaborativefilteringusingtensorflow_b200/dist.py):
request routing by item % P, row fetch, gradient return -- the host-side logic of SURVEY 8(e); the CUDA kernels plug in
between fetch and push on a GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, ld, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from collaborativefilteringusingtensorflow_b200.dist import ItemExchange, item_shard_rows
        ex = ItemExchange(world, rank)
        n_local = item_shard_rows(n_items, world, rank)
        assert sum(item_shard_rows(n_items, world, r) for r in range(world)) == n_items
        # shard row j holds global item j * world + rank; encode the global id in the row so fetches can be verified
        gid = torch.arange(n_local, dtype=torch.float32) * world + rank
        shard = gid[:, None] * 10 + torch.arange(ld, dtype=torch.float32)[None, :] * 0.001
        g = torch.Generator().manual_seed(100 + rank)
        for trial in range(3):
            ids = torch.randint(0, n_items, (64, 4), generator=g)
            if trial == 2:
                ids = torch.full((5, 2), int(rank), dtype=torch.int64)      # every occurrence the same item
            plan = ex.plan(ids, n_items if trial != 1 else None)      # sort-free dedupe and the torch.unique fallback
            rows = ex.fetch(plan, shard)
            assert rows.shape == (plan.n_req, ld) and plan.n_req == len(torch.unique(ids))
            got = rows[plan.occ_local.to(torch.int64).reshape(-1)]
            want = ids.reshape(-1).to(torch.float32)[:, None] * 10 + torch.arange(ld, dtype=torch.float32)[None, :] * 0.001
            assert torch.allclose(got, want), 'fetched rows do not match the requested global ids'
            # gradient return: rank r sends (1 + r) * id for every row it fetched; owners must see one row per requester
            grads = plan.req_global.to(torch.float32)[:, None].repeat(1, ld) * (1 + rank)
            recv = ex.push(plan, grads)
            assert recv.shape[0] == plan.recv_local_rows.numel()
            owned_global = plan.recv_local_rows.to(torch.float32) * world + rank
            ratio = recv[:, 0] / torch.clamp(owned_global, min=1)
            ok = (owned_global == 0) | torch.isin(ratio, torch.arange(1, world + 1, dtype=torch.float32))
            assert bool(ok.all()), 'a returned gradient row does not belong to the row it is aligned with'
            # per-row sums at the owner == sum over requesters
            summed = torch.zeros(n_local).index_add_(0, plan.recv_local_rows.to(torch.int64), recv[:, 0])
            req_mask = torch.zeros(n_items)
            req_mask[torch.unique(ids)] = 1 + rank
            allm = [torch.zeros(n_items) for _ in range(world)]
            dist.all_gather(allm, req_mask)
            expect_global = sum(allm) * torch.arange(n_items, dtype=torch.float32)
            assert torch.allclose(summed, expect_global[rank::world])
            # owner-pull variant: the whole count matrix is gathered, so the owner knows where its segment starts in every
            # requester's gradient buffer; reading those segments in place (here: from an all-gathered copy standing in
            # for NVLink peer memory) must give exactly what the push would have delivered
            plan2 = ex.plan_exchange(ex.plan_local(ids, n_items), all_counts=True)
            assert plan2.send_counts == plan.send_counts and plan2.recv_counts == plan.recv_counts
            assert torch.equal(plan2.recv_local_rows, plan.recv_local_rows)
            cap = 64 * 4
            mine = torch.zeros(cap, ld)
            mine[:plan.n_req] = grads
            every = [torch.zeros(cap, ld) for _ in range(world)]
            dist.all_gather(every, mine)
            pulled = torch.cat([every[r][plan2.peer_offsets[r]:plan2.peer_offsets[r] + plan2.recv_counts[r]] for r in range(world)])
            assert torch.equal(pulled, recv), 'segments read in place differ from the pushed rows'
        q.put((rank, 'ok'))
    except Exception as e:      # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_item_exchange_over_gloo(world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1001, 8, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d: %s' % (rank, msg)


def test_single_rank_exchange_is_identity():
    from collaborativefilteringusingtensorflow_b200.dist import ItemExchange
    ex = ItemExchange(1, 0)
    shard = torch.arange(50, dtype=torch.float32)[:, None].repeat(1, 4)
    ids = torch.tensor([[7, 3], [3, 49]])
    plan = ex.plan(ids)
    rows = ex.fetch(plan, shard)
    assert torch.equal(rows[plan.occ_local.to(torch.int64)][..., 0], ids.to(torch.float32))
    assert torch.equal(ex.push(plan, rows), rows)
