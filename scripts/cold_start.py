"""cold-start costs: context creation, first rule generation, first README-demo call, second context"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectralkernels_jl_b200 as sk
import torch
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
t0 = time.perf_counter(); eng = sk.Session(0); t1 = time.perf_counter()
eng.rule_set(4096, 16, 0.0); t2 = time.perf_counter()
eng.rule_set(4096, 16, -0.5); t3 = time.perf_counter()
rs = 10 ** np.linspace(-6, 0, 1000)
cfg = sk.AdaptiveKernelConfig(sk.Matern(1.0, 1.0, 1.5), engine=eng)
sk.kernel_values(cfg, rs, k0=np.pi / 2); t4 = time.perf_counter()
sk.kernel_values(cfg, rs, k0=np.pi / 2); t5 = time.perf_counter()
e2 = sk.Session(0); t6 = time.perf_counter()
e2.rule_set(4096, 16, 0.0); t7 = time.perf_counter()
print(f"first context {1e3*(t1-t0):.1f} ms; Legendre rules (4096, 8192) {1e3*(t2-t1):.1f} ms; + Jacobi(0,-0.5) rules {1e3*(t3-t2):.1f} ms; "
      f"first README call {1e3*(t4-t3):.2f} ms; second {1e3*(t5-t4):.2f} ms; second context {1e3*(t6-t5):.1f} ms; its rules {1e3*(t7-t6):.1f} ms")
