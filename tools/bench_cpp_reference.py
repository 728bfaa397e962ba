"""CPU baseline (3) of SURVEY.md 8d: the reference's own compiled solver (hpc/solver.cpp, built where it lies by
oracle/Makefile into oracle/_ref/libscpn_solver.so), `run_step` with 200 RB-SOR sweeps on the bench problem,
one thread (the reference build is single-threaded).  Runs on any host; no GPU involved.

    make -C oracle ref && python tools/bench_cpp_reference.py [out.json]
"""
import ctypes, json, os, statistics, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "oracle", "_ref", "libscpn_solver.so")
lib = ctypes.CDLL(so)
dp = ctypes.POINTER(ctypes.c_double)
lib.create_solver.restype = ctypes.c_void_p
lib.create_solver.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4
lib.run_step.argtypes = [ctypes.c_void_p, dp, dp, ctypes.c_int, ctypes.c_int]
lib.destroy_solver.argtypes = [ctypes.c_void_p]

rows = []
for n in (129, 257, 513):
    R = np.linspace(4.0, 8.0, n); Z = np.linspace(-4.0, 4.0, n)
    rr, zz = np.meshgrid(R, Z)
    j = np.ascontiguousarray(np.exp(-((rr - 6.0) ** 2 + zz ** 2) / 0.5))   # bench_gpu_gs_solver._problem source shape
    psi = np.zeros((n, n))
    times = []
    for rep in range(3):
        h = lib.create_solver(n, n, 4.0, 8.0, -4.0, 4.0)
        t0 = time.perf_counter()
        lib.run_step(h, j.ctypes.data_as(dp), psi.ctypes.data_as(dp), n * n, 200)
        times.append(time.perf_counter() - t0)
        lib.destroy_solver(h)
    med = statistics.median(times)
    rows.append({"grid": f"{n}x{n}", "sweeps": 200, "median_s": med, "mlups": 200 * (n - 2) ** 2 / med / 1e6})
out = {"kernel": "libscpn_solver.so run_step (reference hpc/solver.cpp, g++ -O3 -march=native)", "threads": 1,
       "host": os.uname().nodename, "cpu_count": os.cpu_count(), "grids": rows}
text = json.dumps(out, indent=1)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text + "\n")
print(text)
