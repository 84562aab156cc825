"""Where does one objective-only evaluation (a line-search probe) spend its time at C1 size?"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gpuvmem_b200 import Engine, synth, host

cfg = sys.argv[1] if len(sys.argv) > 1 else "c1"
p = synth.config_c1() if cfg == "c1" else synth.config_c2()
host.set_quiet(True)
s = host.Session(p, args="-z 0.001 -Z 0.01 -t 3", fi_spec="Chi2:-1:0:0,Entropy:0:0:0")
s.set_iteration(1)
def t(fn, n=300):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print("calc_function (Chi2+Entropy) us:", round(t(s.calc_function), 1))
e = Engine.from_problem(p)
I = torch.from_numpy(e.initial_image()).cuda()
out = torch.zeros(1, dtype=torch.float64, device="cuda")
print("gvm_chi2 (sync) us:", round(t(lambda: e.chi2(I)), 1))
print("gvm_chi2_async us (no sync in loop):", round(t(lambda: e.chi2_async(I, False, out)), 1))
print("gvm_prior_value us:", round(t(lambda: e.prior_value(0, I)), 1))
xt = torch.empty_like(I)
print("evaluate_xt us:", round(t(lambda: e.vec_evaluate_xt(xt, I, I, 0.5)), 1))
print("launches per chi2:", (lambda a: (e.chi2(I), e.launch_count() - a)[1])(e.launch_count()))
