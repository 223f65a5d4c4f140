"""Multi-GPU parity + timing check, run under torchrun (one rank per GPU, NCCL):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multi_gpu_check.py
Every rank also solves the problem alone; the sharded solves must reproduce it (PTR / AutoPTR to rounding, IAI bit for bit
with identical total evaluation counts).  Prints one JSON line per case from rank 0."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
import autobz_b200 as ab

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = ab.default_context(local)
shard = ab.Shard(rank, world, ab.torch_allreduce(torch.device("cuda", local)))
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
d = np.load(os.path.join(ROOT, "tests", "golden", "svo_hr.npz"))
Hs, los, A = np.asfortranarray(d["H_R"]), tuple(int(x) for x in d["lo"]), d["A"]
fs = ab.FourierSeries(Hs, period=1.0, lo=los, norb=3)
ibz = ab.load_bz(ab.CubicSymIBZ(), A)
ok = True


def timed(fn):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); dist.barrier()
    return r, time.perf_counter() - t


def emit(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


# C3: IAI, outermost panel nodes dealt to the ranks
for eta, w in ((1e-2, 12.0), (1e-4, 12.975161)):
    f = ab.FourierIntegrand(ab.dos_integrand, fs, eta)
    prob = ab.IntegralProblem(f, ibz, w)
    one, t1 = timed(lambda: ab.solve(prob, ab.EvalCounter(ab.IAI()), abstol=1e-3))
    many, tn = timed(lambda: ab.solve(prob, ab.EvalCounter(ab.IAI()), abstol=1e-3, shard=shard))
    good = (one.u == many.u) and (one.numevals == many.numevals)
    ok = ok and good
    emit(case=f"C3 IAI eta={eta} omega={w}", ranks=world, identical=bool(good), numevals=many.numevals, s_1rank=t1, s_sharded=tn, speedup=t1 / tn)

# C2: AutoPTR on the IBZ, k3 planes round-robin over the ranks, one allreduce per rule
f = ab.FourierIntegrand(ab.gloc_trace_integrand, fs, eta=1e-2)
prob = ab.IntegralProblem(f, ibz, {"omega": 12.5})
alg = ab.EvalCounter(ab.AutoPTR(a=1e-2, nmin=50, nmax=1000))
one, t1 = timed(lambda: ab.solve(prob, alg, abstol=1e-3))
many, tn = timed(lambda: ab.solve(prob, alg, abstol=1e-3, shard=shard))
good = abs(one.u - many.u) <= 1e-12 * abs(one.u) and one.numevals == many.numevals
ok = ok and good
emit(case="C2 AutoPTR(a=eta=1e-2) CubicSymIBZ omega=12.5", ranks=world, rel_diff=abs(one.u - many.u) / abs(one.u), numevals=many.numevals, s_1rank=t1,
     s_sharded=tn, speedup=t1 / tn)

# C5: norb = 64 band energy on the IBZ, AutoPTR 48 -> 96 -> 144
H5, lo5 = ab.synthetic.wannier_hamiltonian(64, 4, cubic=True)
f5 = ab.FourierIntegrand(ab.EigenIntegrand("fermi_energy"), ab.FourierSeries(H5, period=1.0, lo=lo5, norb=64), 0.0, 0.5)
cub = ab.load_bz(ab.CubicSymIBZ(), 2 * np.pi * np.eye(3))
alg5 = ab.EvalCounter(ab.AutoPTR(a=1.0, nmin=48, dn=48.0))
ab.solve(ab.IntegralProblem(f5, cub), ab.PTR(npt=24))
one, t1 = timed(lambda: ab.solve(ab.IntegralProblem(f5, cub), alg5, reltol=1e-6))
many, tn = timed(lambda: ab.solve(ab.IntegralProblem(f5, cub), alg5, reltol=1e-6, shard=shard))
good = abs(one.u - many.u) <= 1e-12 * abs(one.u) and one.numevals == many.numevals
ok = ok and good
emit(case="C5 norb=64 band energy CubicSymIBZ AutoPTR 48->96->144", ranks=world, rel_diff=abs(one.u - many.u) / abs(one.u), numevals=many.numevals,
     kpoints_per_s_1rank=one.numevals / t1, kpoints_per_s_sharded=many.numevals / tn, speedup=t1 / tn)
# the library's own NCCL path (abz_comm_init / abz_allreduce_sum through dlopen): unique id from rank 0 over torch.distributed,
# then a raw allreduce and a sharded IAI solve whose exchange is NCCL inside the library (exchange callback = NULL)
from autobz_b200 import _lib as L
uid = [L.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
v = np.arange(4, dtype=np.float64) + rank
ctx.allreduce_sum(v)
good = bool(np.array_equal(v, world * np.arange(4) + world * (world - 1) / 2))
f = ab.FourierIntegrand(ab.dos_integrand, fs, 1e-2)
nest = L.DeviceNest(ctx, fs.device(ctx), 3, 64, 2048)
jb = abs(np.linalg.det(ibz.B)) * 48
args = (1, ibz.lims.a, None, L.F_RESOLVENT_TRACE, 1, complex(12.0, 1e-2), None, None, 1e-3 / jb, 0.0, 2 ** 62)
I1, E1, ne1, _, _ = nest.iai_solve(*args)
In, En, nen, _, _ = nest.iai_solve(*args, rank=rank, nranks=world, allreduce=None)
good = good and I1 == In and E1 == En and ne1 == nen
ok = ok and good
emit(case="library NCCL path: abz_allreduce_sum + abz_iai_solve_sharded(exchange=NULL)", ranks=world, identical=bool(good), numevals=nen,
     exchanges=nest.last_exchanges)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(int(flag.item() != 0))
