import sys, time, ctypes, numpy as np, torch
sys.path.insert(0, "/root/repo/nbody-deep-sim_b200"); sys.path.insert(0, "/root/repo")
from galaxify import galaxies, simulation, _native
L = _native.lib()
for packed in (0, 1):
    tf = ctypes.c_double()
    _native.call("nbody_probe_fp32_peak", 0, packed, ctypes.byref(tf)); print("fp32 peak probe packed=%d: %.2f TFLOP/s" % (packed, tf.value))
for n in (16384, 65536, 262144, 1048576):
    pos, vel, mass = galaxies.generate_plummer(n_bodies=n, total_mass=1.0, scale_radius=1.0, g_const=1.0, seed=1)
    sim = simulation.LeapFrogSimulator(positions=pos, velocities=vel, masses=mass, g_const=1.0, softening=0.01, dt=1e-3, calc_energy=False)
    torch.cuda.synchronize()
    reps = 3 if n >= 262144 else 20
    sim.compute_accelerations(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): sim.compute_accelerations()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("N=%d accel %.3f ms  %.3e interactions/s  (%.1f%% of 74.4 TF @20flop)" % (n, ms, n*n/(ms*1e-3), 20*n*n/(ms*1e-3)/74.4e12*100))
    st = np.zeros(5, dtype=np.float32)
    sim._integrate(5, 1, None, None, st); print("   leapfrog step ms:", st)
