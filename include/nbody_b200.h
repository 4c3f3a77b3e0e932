/* nbody_b200.h — C ABI of the B200-native direct-summation N-body engine.
 *
 * This is the drop-in boundary for the hot path of bikuta6/nbody-deep-sim's `galaxify` simulator. The reference
 * has no FFI of its own (it is pure Python on torch); each entry point below cites the reference interface it
 * replaces, relative to the reference root. The Python host side (nbody-deep-sim_b200/galaxify/) binds these
 * symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - All arithmetic is FP32 (the reference converts to torch.float32 at simulation.py:58-65).
 *   - State arrays use the reference's layout: positions/velocities/accelerations are (n,3) contiguous floats,
 *     masses are (n,) floats.
 *   - "_f32" entry points take DEVICE pointers and a CUDA stream (cudaStream_t passed as void*; NULL = default
 *     stream) and never synchronise unless stated. "_host_f32" entry points take HOST pointers, do their own
 *     host<->device copies on an internal stream and return after synchronising.
 *   - `dt` and `half_dt` are passed already rounded to FP32 by the caller, as torch does for `0.5 * dt * tensor`
 *     (simulation.py:164): half_dt = (float)(0.5 * dt_double), dt = (float)dt_double.
 *   - `eps2` is (float)(softening * softening) computed in double (simulation.py:82).
 *   - Every function returns NBODY_OK (0) or a negative NBODY_ERR_* code and never throws across the ABI.
 *     nbody_last_error() gives a per-thread message for the last failure.
 *   - The library is safe to call from one host thread per GPU.
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBODY_OK 0
#define NBODY_ERR_INVALID_ARGUMENT (-1) /* null pointer, n < 1, negative count, bad enum */
#define NBODY_ERR_WORKSPACE (-2)        /* workspace too small for this problem */
#define NBODY_ERR_CUDA (-3)             /* a CUDA runtime call or launch failed */
#define NBODY_ERR_NO_DEVICE (-4)        /* no CUDA device / not an sm_100 device */
#define NBODY_ERR_UNSUPPORTED (-5)      /* shape outside what the entry point supports */

#define NBODY_INTEGRATOR_LEAPFROG 1 /* simulation.py:153-170 */
#define NBODY_INTEGRATOR_EULER 2    /* simulation.py:173-187 */

/* ABI version (major*100 + minor). */
int nbody_version(void);
const char* nbody_status_string(int status);
const char* nbody_last_error(void);

/* Number of kernels this library has launched on behalf of the calling process (all threads). */
uint64_t nbody_launch_count(void);

/* ---------------------------------------------------------------- workspace ---------------------------------- */

/* Bytes of device scratch a stepping/acceleration call needs for `n_local` i-bodies against `n_total` j-bodies
 * (single GPU: n_local == n_total). The caller allocates it (e.g. torch.empty(uint8)); contents are scratch and
 * need no initialisation. */
size_t nbody_workspace_bytes(int n_local, int n_total);

/* The launch plan the force kernel will use for `n_local` i-bodies against a j range of `j_len` bodies: kernel shape
 * (1 = 512 threads x 4 i-bodies, 0 = 256 threads x 2 i-bodies), number of i-tiles and of j-splits. A pure function of
 * its arguments (planned for the 148 SMs of a B200); exposed for tests and for sizing expectations. */
int nbody_plan_f32(int n_local, int j_len, int* shape_large, int* i_tiles, int* splits);

/* ---------------------------------------------------------------- single-system path ------------------------- */

/* Replaces BaseSimulator.compute_accelerations (simulation.py:71-89).
 * acc[i] = fl32(g) * sum_j m_j (r_j - r_i) (|r_j - r_i|^2 + eps2)^(-3/2), self term exactly zero. */
int nbody_accel_f32(const float* pos, const float* mass, float* acc, int n, float g, float eps2, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Replaces `steps` iterations of BaseSimulator.run's loop body (simulation.py:126-146) with
 * LeapFrogSimulator.step (:153-170) or EulerSimulator.step (:173-187) fused into the force kernel's epilogue.
 *
 * In/out: pos, vel, acc hold the state before the call and the state after `steps` steps on return
 * (acc must hold a(pos) on entry for leapfrog, which BaseSimulator.__init__ guarantees, simulation.py:69).
 * traj (optional, may be NULL): receives every `record_every`-th state, step s (0-based) is recorded when
 *   (s + 1) % record_every == 0, into slot (s + 1) / record_every - 1; layout [slot][3][n][3] floats with planes
 *   positions, velocities, accelerations — the three tensors of one SimulationState (simulation.py:135-145).
 * energies (optional, may be NULL): device doubles [slot][2] = (u_energy, k_energy) of each recorded state,
 *   computed as compute_energies does (simulation.py:91-115), with softening `eps` entering as |r| + eps.
 * step_ms (optional HOST pointer, may be NULL): per-step device time in milliseconds from CUDA events. Passing it
 *   makes the call synchronise the stream before returning. */
int nbody_integrate_f32(int integrator, float* pos, float* vel, float* acc, const float* mass, int n, float g,
                        float eps2, float eps, float dt, float half_dt, int steps, int record_every, float* traj,
                        double* energies, float* step_ms, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces BaseSimulator.compute_energies (simulation.py:91-115). out_uk: device doubles (u_energy, k_energy). */
int nbody_energies_f32(const float* pos, const float* vel, const float* mass, int n, float g, float eps,
                       double* out_uk, void* workspace, size_t workspace_bytes, void* stream);

/* Total linear and angular momentum of a state: out = 6 device doubles (px, py, pz, lx, ly, lz) with p = sum m v and
 * l = sum m (x cross v), FP64 sums in a fixed order. `pos` may be NULL (then l = 0). The reference has no momentum
 * function; its momentum DRIFT is one of the parity metrics, evaluated on the device here. */
int nbody_momentum_f32(const float* pos, const float* vel, const float* mass, int n, double* out, void* stream);

/* ---------------------------------------------------------------- rollout integrator (force slot left open) -- */

/* The kick-drift-kick update that the surrogate-model rollout wraps around `model.predict`
 * (trainer.py:217-226 Trainer.step, gnn.py:223-232 GraphModel.step), i.e. LeapFrogSimulator.step
 * (simulation.py:164-170) with the accelerations supplied by the caller:
 *   nbody_kick_drift_f32:  vel_out = vel + half_dt*acc ; pos_out = pos + dt*vel_out      (trainer.py:219-221)
 *   ... caller evaluates acc_new = predict(pos_out, ...) ...                             (trainer.py:223)
 *   nbody_kick_f32:        vel_out = vel + half_dt*acc_new                               (trainer.py:225)
 * All arrays are (n,3) device floats; outputs may alias the matching inputs (in-place) or be distinct (the reference
 * returns new tensors). Each update is a separately rounded multiply and add, as torch evaluates it. n = 0 is a no-op. */
int nbody_kick_drift_f32(const float* pos, const float* vel, const float* acc, float* pos_out, float* vel_out, int n,
                         float dt, float half_dt, void* stream);
int nbody_kick_f32(const float* vel, const float* acc, float* vel_out, int n, float half_dt, void* stream);

/* ---------------------------------------------------------------- sharded (multi-GPU) path ------------------- */

/* i-sharded building blocks: every rank holds the full body array (x,y,z,m float4 x n_total) and the state of
 * its own slice [i_begin, i_begin + n_local). There is no reference counterpart (the reference is single
 * device); these compute exactly what the single-system path computes for that slice.
 *
 * nbody_shard_prepare_f32: writes this rank's slice of `bodies` from (pos, mass) and, for leapfrog, performs the
 *   opening half-kick + drift (simulation.py:164-166) into pos / vhalf first. The caller then all-gathers bodies.
 * nbody_shard_force_f32: accumulates the j-range [j_begin, j_end), plus an optional second range
 *   [j2_begin, j2_end) (pass j2_begin >= j2_end for none), for the local i-bodies. A step may be split over several
 *   calls (own slice first, the rest once gathered); `part` / `n_parts` identify them and the last call to complete
 *   runs the integrator epilogue and writes this rank's slice of `bodies_next`. Bodies with zero mass contribute
 *   nothing, so ranges may include padding entries as long as those sit away from every real body. */
size_t nbody_shard_workspace_bytes(int n_local, int n_total, int n_parts);
int nbody_shard_prepare_f32(int integrator, float* pos, const float* vel, const float* acc, const float* mass,
                            float* vhalf, float* bodies, int i_begin, int n_local, float dt, float half_dt,
                            void* stream);
int nbody_shard_force_f32(int integrator, const float* bodies, float* bodies_next, int n_total, int i_begin,
                          int n_local, int j_begin, int j_end, int j2_begin, int j2_end, int part, int n_parts,
                          float* pos, float* vel, float* acc, float* vhalf, float g, float eps2, float dt,
                          float half_dt, int do_next, void* workspace, size_t workspace_bytes, void* stream);

/* This rank's share of compute_energies (simulation.py:91-115): u = -G sum_{i local} m_i sum_{j > i} m_j/(|r_ij|+eps)
 * over the full body array (pad slots must hold zero masses), k over the local velocities. The caller sums the two
 * doubles over ranks. Workspace: nbody_shard_workspace_bytes(n_local, n_total, 1) suffices. */
int nbody_shard_energies_f32(const float* bodies, const float* vel, int n_total, int i_begin, int n_local, float g,
                             float eps, double* out_uk, void* workspace, size_t workspace_bytes, void* stream);

/* Pair (Newton's-third-law) building blocks of the sharded step, for systems of at least nbody_pair_min_bodies()
 * bodies: every unordered pair of bodies is evaluated once over ALL ranks. Rank `my_slot` evaluates the triangle of
 * its own slot, the full rectangles against the next floor((P-1)/2) slots and, for even P, half of the rectangle
 * against the opposite slot; forces and reactions are accumulated (RED.ADD.F64) into `acc64`, an array of
 * n_slots*slot_size*3 doubles laid out like the body array, which therefore holds partial sums for bodies of OTHER
 * slots too: the caller sums the arrays over ranks slot by slot (a reduce-scatter) and hands its own slot's sums to
 * nbody_shard_pair_finish_f32, which applies fl32(G) and the integrator epilogue exactly as nbody_shard_force_f32's
 * last part does, writes the rank's slice of bodies_next, and clears acc_own plus `acc_clear` (pass the full acc64
 * there when it is a different buffer from acc_own; NULL/0 otherwise) for the next step.
 *   plan : once per simulator (and again if n or the layout changes); split_phases != 0 keeps the own-slot triangle
 *          (phase 0, needs no remote data: run it while the all-gather is in flight) apart from the cross-slot
 *          rectangles (phase 1); split_phases == 0 puts everything into phase 0.
 *   force: one launch per phase, persistent CTAs pulling (I-tile, J-run) items from a counter in the workspace.
 * `n` is the number of real bodies; slot s holds bodies [s*slot_size, min(n, (s+1)*slot_size)). The workspace content
 * (item lists, counters) must be preserved between plan, force and finish. No reference counterpart. */
int nbody_pair_min_bodies(void);
size_t nbody_shard_pair_workspace_bytes(int n_slots, int slot_size);
/* The part of the interaction matrix rank `my_slot` evaluates, as up to `max_blocks` blocks of 5 ints
 * (i_lo, i_hi, j_lo, j_hi, triangle) in global body indices: block 0 is the triangle of the own slot (unordered pairs
 * within [i_lo, i_hi)), the others are rectangles (every i in [i_lo, i_hi) with every j in [j_lo, j_hi)). Over all ranks
 * the blocks cover every unordered pair of the n bodies exactly once. Pure host function (no device needed); returns the
 * number of blocks or a negative NBODY_ERR_* code. */
int nbody_shard_pair_blocks(int n, int n_slots, int slot_size, int my_slot, int* blocks, int max_blocks);
int nbody_shard_pair_plan_f32(int n, int n_slots, int slot_size, int my_slot, int split_phases, void* workspace,
                              size_t workspace_bytes, void* stream);
int nbody_shard_pair_force_f32(int phase, const float* bodies, int n_slots, int slot_size, float eps2, double* acc64,
                               void* workspace, size_t workspace_bytes, void* stream);
int nbody_shard_pair_finish_f32(int integrator, const float* bodies, float* bodies_next, int i_begin, int n_local,
                                double* acc_own, double* acc_clear, long long acc_clear_count, float* pos, float* vel,
                                float* acc, float* vhalf, float g, float dt, float half_dt, int do_next, int n_slots,
                                int slot_size, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- batched many-small-systems path ------------ */

/* `n_systems` independent systems of `n` bodies each (n <= nbody_batched_max_n()), all stepped `steps` times
 * inside one persistent kernel (one CTA per system, bodies resident in shared memory). Arrays carry a leading
 * system dimension: pos/vel/acc (n_systems,n,3), mass (n_systems,n). Each system evolves exactly as
 * nbody_integrate_f32 would evolve it alone, up to FP32 summation order. traj layout:
 * [slot][3][n_systems][n][3]. No reference counterpart: s01-dataset-generation.py:130-214 runs scenes one by one. */
int nbody_batched_max_n(void);
int nbody_batched_integrate_f32(int integrator, float* pos, float* vel, float* acc, const float* mass,
                                int n_systems, int n, float g, float eps2, float dt, float half_dt, int steps,
                                int record_every, float* traj, void* stream);
int nbody_batched_accel_f32(const float* pos, const float* mass, float* acc, int n_systems, int n, float g,
                            float eps2, void* stream);
/* compute_energies (simulation.py:91-115) for every recorded state of a trajectory buffer written by
 * nbody_batched_integrate_f32 (or by nbody_integrate_f32, which is the n_systems == 1 layout), all states in
 * parallel. out: device doubles [slot][n_systems][2] = (u_energy, k_energy). n <= nbody_batched_max_n(),
 * n_slots <= 65535 per call. */
int nbody_traj_energies_f32(const float* traj, const float* mass, int n_slots, int n_systems, int n, float g,
                            float eps, double* out, void* stream);

/* ---------------------------------------------------------------- host-buffer path --------------------------- */

/* Same operations with HOST buffers: copies inputs to the device, runs, copies results back, synchronises.
 * `device` is the CUDA device ordinal. Scratch is cached per device inside the library between calls.
 * h2d_bytes/d2h_bytes (optional) receive the bytes copied in each direction. */
int nbody_accel_host_f32(const float* pos, const float* mass, float* acc, int n, float g, float eps2, int device,
                         uint64_t* h2d_bytes, uint64_t* d2h_bytes);
int nbody_integrate_host_f32(int integrator, float* pos, float* vel, float* acc, const float* mass, int n, float g,
                             float eps2, float eps, float dt, float half_dt, int steps, int record_every,
                             float* traj, double* energies, float* step_ms, int device, uint64_t* h2d_bytes,
                             uint64_t* d2h_bytes);
/* Frees the cached scratch of nbody_*_host_f32 on every device. */
int nbody_host_cache_release(void);

/* ---------------------------------------------------------------- measurement -------------------------------- */

/* Measures the FP32 peak of `device` with a register-resident FFMA microkernel (packed = 0: scalar FFMA;
 * packed = 1: FFMA2). Returns TFLOP/s counting 2 flops per FMA lane. Used as the roofline denominator because
 * MEASURED_PEAKS.json carries no FP32 entry. Synchronises. */
int nbody_probe_fp32_peak(int device, int packed, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
