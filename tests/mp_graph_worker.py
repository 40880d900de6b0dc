"""One rank of the multi-GPU parity test (launched by torch.distributed.run from test_gpu_determinism.py, or by hand:
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mp_graph_worker.py).  Every rank
builds its shard, the library gathers; rank 0 also builds the graph alone and compares bit for bit."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--space", default="PCA")
    ap.add_argument("--samples", type=int, default=6)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import conos_b200
    from conos_b200.panel import make_panel
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = conos_b200.Context(local)
    cells = [900, 700, 1100, 600, 800, 1000, 750, 650][:args.samples]
    panel = make_panel(11, args.samples, cells, n_genes=400, n_types=8, n_pcs=20, use_torch=False)
    ncomps = {"PCA": 20, "CPCA": 12, "CCA": 10}[args.space]
    kw = dict(k=10, k_self=5, space=args.space, ncomps=ncomps, n_odgenes=400)
    con = conos_b200.Conos(panel, ctx=ctx)
    g = con.buildGraph(world=(rank, world, None), **kw)
    ok = True
    if rank == 0:
        solo = conos_b200.Conos(panel, ctx=ctx)
        g1 = solo.buildGraph(**kw)
        for name in ("vertices", "edge_from", "edge_to", "weight", "type", "adj_p", "adj_i", "adj_x"):
            same = np.array_equal(getattr(g, name), getattr(g1, name))
            if not same:
                print(f"MP_GRAPH_DIFF {name}", flush=True)
            ok = ok and same
        print(f"rank 0: {g.vcount()} vertices, {g.ecount()} edges; single rank {g1.ecount()} edges", flush=True)
        solo.close()
    else:
        assert g is None
    con.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    if rank == 0 and int(flag.item()) == 1:
        print("MP_GRAPH_EQUAL", flush=True)
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
