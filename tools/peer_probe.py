"""2-GPU probe (torchrun): where does the time of the peer-exchange step go?
Times, per step and as the max over ranks: the full step, the two sync kernels alone, the merge
with and without peer stores, the gather.  cfg2, CUDA graphs of 20 iterations."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS, Workset, make_plan, make_runner, timed  # noqa: E402
from dynamicscaler_b200 import _lib  # noqa: E402
from dynamicscaler_b200.step import PanoramaStepper  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = WORKLOADS["cfg2_sphere_128x256"]
st = PanoramaStepper(wl["H"], wl["W"], wl["F"], device=dev, rank=rank, world=world)
st.use_multicast = os.environ.get("DS_NO_MULTICAST") is None
gen = torch.Generator(device=dev).manual_seed(rank)
gp = torch.Generator(device=dev).manual_seed(99)
res = {"multicast": os.environ.get("DS_NO_MULTICAST") is None, "world": world}
side = torch.cuda.Stream()
for mode in ("nearest", "bilinear"):
    plans = [make_plan(st, wl, mode, i) for i in range(10)]
    j0, nj = st.local_jobs(plans[0])
    sets = [Workset(wl, nj, dev, gen, st, True, gp) for _ in range(4)]
    combos = [(i % 10, i % 4) for i in range(20)]
    G, M = _lib.DS_PEER_AFTER_GATHER, _lib.DS_PEER_AFTER_MERGE

    def seq(kind):
        fns = []
        for off, s in combos:
            ws, plan = sets[s], plans[off]
            state = ws.state
            pms = state.pms
            mask = state.mask_plane

            def full(ws=ws, plan=plan, state=state):
                st.gather_all(state, plan, out=ws.wins)
                st.merge_all(state, [ws.x_prev, ws.x0], plan)

            def syncs(state=state):
                st._peer_sync(state, G)
                st._peer_sync(state, M)

            def merge_peer(ws=ws, plan=plan, state=state, pms=pms, mask=mask):
                st._pm_merge_into(pms, [ws.x_prev, ws.x0], mask, _lib.DS_MASK_PLANE, plan, j0, nj, state.peer_table)
                syncs(state)

            def merge_local(ws=ws, plan=plan, state=state, pms=pms, mask=mask):
                st._pm_merge_into(pms, [ws.x_prev, ws.x0], mask, _lib.DS_MASK_PLANE, plan, j0, nj, None)
                syncs(state)

            def gather(ws=ws, plan=plan, state=state):
                st._pm_gather(state.pms[0], plan, ws.wins, j0, nj, None)
                syncs(state)

            fns.append({"full": full, "syncs": syncs, "merge_peer": merge_peer, "merge_local": merge_local,
                        "gather": gather}[kind])
        return fns

    for kind in ("full", "syncs", "merge_peer", "merge_local", "gather"):
        run = make_runner(seq(kind), 200, True, side)
        run()
        ms = timed(run, True)
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"{mode}/{kind}_us"] = round(float(t.item()) / 200 * 1e3, 2)
    del sets
if rank == 0:
    print(json.dumps(res))
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
