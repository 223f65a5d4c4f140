"""Quick device-vs-oracle check of the low-level C ABI (development aid; the real tests live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import numpy as np
import orc
import autobz_b200
from autobz_b200 import _lib as L, synthetic

ctx = L.Context(0)
ok = True
def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
def report(name, err, tol=1e-11):
    global ok
    flag = "OK " if err <= tol else "BAD"
    if err > tol: ok = False
    print(f"{flag} {name}: rel err {err:.3e}", flush=True)

# 1. C1
c, lo = synthetic.integer_lattice(3)
S = L.DeviceSeries(ctx, c, lo, (1.0,)*3); So = orc.Series(c.astype(complex), lo)
z = np.array([0.1j, 0.5 + 0.1j])
R = L.DeviceRule(ctx, S, 16)
report("C1 fused N=16", rel(R.resolvent_sum(z, scale=1/16**3), orc.ptr_sum(So, 16, z)))
H, k, w = R.copy_out()
report("C1 H copy_out", rel(H.reshape(1,1,16,16,16, order="F"), orc.grid_eval_full(So, 16)))
R.materialize()
report("C1 materialised", rel(R.resolvent_sum(z, scale=1/16**3), orc.ptr_sum(So, 16, z)))
report("C1 trace H", rel(R.resolvent_sum(None, scale=1/16**3, fkind=L.F_TRACE_H), orc.ptr_sum(So, 16, None, fkind=1)), 1e-9)
R64 = L.DeviceRule(ctx, S, 64)
g = R64.resolvent_sum(z, scale=1/64**3)
print("C1 N=64", g, "expect -2.361629003144814i, 1.448114087711081-1.4016191114904277i")

# 2. n=3
for n, rmax, N in [(3, 2, 12), (2, 1, 9), (5, 2, 10), (8, 1, 8), (32, 1, 6), (17, 1, 5), (64, 1, 4)]:
    Hc, lo = synthetic.wannier_hamiltonian(n, rmax)
    S = L.DeviceSeries(ctx, Hc, lo, (1.0,)*3); So = orc.Series(Hc, lo)
    z = np.array([0.3 + 0.05j, -0.7 + 0.2j, 1.5 + 0.01j])
    rng = np.random.default_rng(n)
    sig = 0.1 * (rng.standard_normal((n, n, 3)) + 1j * rng.standard_normal((n, n, 3)))
    R = L.DeviceRule(ctx, S, N)
    t = time.time(); ref = orc.ptr_sum(So, N, z); t = time.time() - t
    report(f"n={n} N={N} streamed", rel(R.resolvent_sum(z, scale=1/N**3), ref))
    report(f"n={n} N={N} streamed+sigma", rel(R.resolvent_sum(z, sigma=sig, scale=1/N**3), orc.ptr_sum(So, N, z, sigma=sig)))
    H, k, w = R.copy_out()
    report(f"n={n} H copy_out", rel(H.reshape(n, n, N, N, N, order="F"), orc.grid_eval_full(So, N)), 1e-12)
    R.materialize()
    report(f"n={n} N={N} materialised", rel(R.resolvent_sum(z, scale=1/N**3), ref))
    # slab
    Rs = L.DeviceRule(ctx, S, N, k3_lo=1, k3_hi=3)
    report(f"n={n} slab", rel(Rs.resolvent_sum(z, scale=1/N**3), orc.ptr_sum(So, N, z, k3_lo=1, k3_hi=3)))
    # eig
    ev = R.eigvals()
    Hm = np.moveaxis(H, 2, 0)
    ref_ev = np.linalg.eigvalsh(Hm)
    report(f"n={n} eigvals vs LAPACK", rel(ev, ref_ev), 1e-11)
    report(f"n={n} eig_sum fermi", abs(R.eig_sum(L.EIG_FERMI_ENERGY, (0.1, 0.3), 1/N**3) - orc.ptr_eig_sum(So, N, 1, (0.1, 0.3), scale=1/N**3)[0]), 1e-11)
    # points
    kp = rng.random((7, 3))
    report(f"n={n} points_eval", rel(S.eval_points(kp), orc.eval_points(So, kp)), 1e-12)
    report(f"n={n} points_resolvent", rel(S.points_resolvent(kp, z), orc.resolvent_trace_batch(orc.eval_points(So, kp), z)))

# 3. symmetric rules (cubic-symmetric series)
import itertools
syms = []
for perm in itertools.permutations(range(3)):
    for sg in itertools.product((1, -1), repeat=3):
        m = np.zeros((3, 3), dtype=np.int32)
        for i in range(3): m[i, perm[i]] = sg[i]
        syms.append(m)
for n, rmax, N in [(3, 2, 12), (3, 2, 15), (6, 1, 10)]:
    Hc, lo = synthetic.wannier_hamiltonian(n, rmax, cubic=True)
    S = L.DeviceSeries(ctx, Hc, lo, (1.0,)*3); So = orc.Series(Hc, lo)
    w_o, nirr_o = orc.symptr_rule(N, syms)
    w_d, nirr_d = ctx.symptr_rule(N, syms)
    print("symptr equal:", np.array_equal(w_o, w_d), nirr_o, nirr_d, int(w_o.sum()), N**3)
    if not np.array_equal(w_o, w_d): ok = False
    z = np.array([0.3 + 0.05j, -0.7 + 0.2j])
    R = L.DeviceRule(ctx, S, N, wsym=w_d)
    full = orc.ptr_sum(So, N, z)
    ref, cnt = orc.symptr_sum(So, N, w_o, z, scale=1/N**3)
    report(f"sym n={n} N={N} oracle sym vs full", rel(ref, full))
    report(f"sym n={n} N={N} device ({len(R)} nodes, oracle {cnt})", rel(R.resolvent_sum(z, scale=1/N**3), ref))
    R.materialize()
    report(f"sym n={n} N={N} device materialised", rel(R.resolvent_sum(z, scale=1/N**3), ref))
    R2 = L.DeviceRule(ctx, S, N, wsym=w_d, k3_lo=1, k3_stride=2)
    R1 = L.DeviceRule(ctx, S, N, wsym=w_d, k3_lo=0, k3_stride=2)
    report(f"sym n={n} scatter shards", rel(R1.resolvent_sum(z, scale=1/N**3) + R2.resolvent_sum(z, scale=1/N**3), ref))
    report(f"sym n={n} eig_sum", abs(R.eig_sum(L.EIG_SUM, scale=1/N**3) - orc.ptr_eig_sum(So, N, 0, (0, 1), wsym=w_o, scale=1/N**3)[0]), 1e-11)

# 4. nest
for n in (3, 6):
    Hc, lo = synthetic.wannier_hamiltonian(n, 2)
    S = L.DeviceSeries(ctx, Hc, lo, (1.0,)*3); So = orc.Series(Hc, lo)
    nest = L.DeviceNest(ctx, S, 3, 4, 8)
    rng = np.random.default_rng(5)
    x3 = rng.random(3); nest.contract3(x3, [0, 1, 3])
    x2 = rng.random(5); par = np.array([0, 1, 3, 3, 0]); nest.contract2(x2, par, [7, 0, 2, 3, 5])
    x1 = rng.random(11); s1 = rng.choice([7, 0, 2, 3, 5], 11)
    m = {7: 0, 0: 1, 2: 2, 3: 3, 5: 4}
    map3 = {0: 0, 1: 1, 3: 2}
    kp = np.array([[x1[i], x2[m[s1[i]]], x3[map3[par[m[s1[i]]]]]] for i in range(11)])
    y = nest.eval(x1, s1, 0.2 + 0.1j)
    report(f"nest n={n}", rel(y, orc.resolvent_trace_batch(orc.eval_points(So, kp), [0.2 + 0.1j])[:, 0]))
print("launches", ctx.launch_count)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
