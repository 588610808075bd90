// flybody_b200.cu -- C ABI (include/flybody_b200.h) + kernel launch sequence of the batched fly
// stepper for sm_100a.
//
// One control step = n_substeps x [ step2 ; step1 ] exactly as dm_control's legacy Physics.step()
// (SURVEY.md App. A), each substep being the staged kernel pipeline
//   step2: k_act -> k_smooth(solve) -> k_ref -> k_solve -> k_finish(qacc, sensors, Euler)
//   step1: k_pos(kinematics, CRB, factor) -> k_col -> k_con -> k_proj(J, Z, A) -> k_vel(RNE, passive)
// One warp per env, all per-env data in one record per env (fb_types.h).
//
// The same translation unit compiles as plain C++ with -DFB_EMU (tests/_emu): phases run in
// nested host loops.  That build exists ONLY so CPU tests can exercise the kernel source; the
// Python package never loads it (flybody_b200/stepper.py refuses to run without the CUDA library).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

#ifndef FB_EMU
#include <cuda_runtime.h>
#endif
#include "fb_solver_reg.h"
#include "fb_render.h"
#include "fb_hfield.h"

#ifdef FB_EMU
typedef int cudaStream_t_;
#endif

// -------------------------------------------------------------------------------------------
// memory / launch abstraction
#ifndef FB_EMU
#define FB_CUDA_OK(call) do { cudaError_t err_ = (call); if (err_ != cudaSuccess) { s->err = std::string(#call) + ": " + cudaGetErrorString(err_); return -2; } } while (0)
static void* dev_alloc(size_t bytes) {   // rare (create / first use): zero-fill and make it visible to every stream
  void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) return nullptr; cudaMemset(p, 0, bytes ? bytes : 4); cudaDeviceSynchronize(); return p; }
static void dev_free(void* p) { cudaFree(p); }
static void h2d(void* dst, const void* src, size_t n) { cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice); cudaDeviceSynchronize(); }
static void d2h(void* dst, const void* src, size_t n) { cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost); }
#else
#define FB_CUDA_OK(call) do { } while (0)
static void* dev_alloc(size_t bytes) { return calloc(bytes ? bytes : 4, 1); }
static void dev_free(void* p) { free(p); }
static void h2d(void* dst, const void* src, size_t n) { memcpy(dst, src, n); }
static void d2h(void* dst, const void* src, size_t n) { memcpy(dst, src, n); }
#endif

#ifndef FB_SPLIT_DEFAULT
#define FB_SPLIT_DEFAULT 1
#endif
#ifndef FB_FUSE_DEFAULT
#define FB_FUSE_DEFAULT 0
#endif
#ifndef FB_OVERLAP_VEL_DEFAULT
#define FB_OVERLAP_VEL_DEFAULT 0
#endif
enum { K_ACT = 0, K_SMOOTH, K_REF, K_SOLVE, K_FINISH, K_POS, K_COL, K_CON, K_PROJ, K_VEL, K_SENS, K_PACK, K_MISC, K_STEP2, K_STEP1, K_STEP, K_NKIND };
static const char* const kKindNames[K_NKIND] = {"act", "smooth", "ref", "solve", "finish", "pos", "col", "con", "proj", "vel", "sens", "pack", "misc", "step2", "step1", "step"};
#ifndef FB_EMU
struct ProfEvent { int kind; cudaEvent_t a, b; };
#endif
struct FbSim {
  DevModel m; DevData d;
  int prof_on; double prof_ms[K_NKIND]; long long prof_n[K_NKIND];
#ifndef FB_EMU
  std::vector<ProfEvent> prof_events;
#endif
  FbModel hm;                       // host copy of scalar fields (pointers invalid after create)
  std::vector<void*> allocs;
  std::vector<void*> obs_scope, task_scope;   // device buffers of the current observation / task program: freed when the program is replaced
  std::vector<int> h_dof_parent, h_dof_Madr, h_body_lastdof, h_geom_bodyid;
  std::vector<double> h_qpos0;
  int device, n_sm; long long launches; float last_ms; std::string err;
  int first_substep; int hold_pending;
  int* rst_ids_dev; float* rst_qpos_dev; float* rst_qvel_dev; int rst_cap;
  const int* act_map_dev; int n_action;
  int cur_e0, cur_n;           // env range of the step kernels being issued (whole batch unless FB_SPLIT)
  int fuse;                    // FB_FUSE: how the stage kernels of a step are grouped into launches (see fb_launch_fused)
  int blob_in_smem;            // copy the sweep program of the triangular solves into shared memory per CTA (default: for batches <= 1024 envs; FB_BLOB=0/1 overrides)
  int* op_step_dev; unsigned char* op_first_dev;
  float* ref_slots; int ref_slot_len;      // per-env reference tables (fb_ref_slots)
  DevHf hf; bool hf_on; int hf_nrow, hf_ncol;      // heightfield collision (fb_hfield_collision)
  float *bank_dev, *bank_hmax_dev, *bank_cmax_dev; int bank_n;      // terrain bank of the device-side vision task (fb_hfield_bank)
  DevEye eye; float* hfield_dev; float* hmax_dev; float* cmax_dev; int hf_nbr, hf_nbc; unsigned char* eye_out; size_t eye_bytes;   // eye cameras (fb_eye_program)
  DevTask task_host;                       // host copy of the device-side task program (its pointers are device pointers)
  float* stage; int* stage_i; size_t stage_cap, stage_icap; unsigned ws_slot;
#ifndef FB_EMU
  cudaStream_t stream; cudaEvent_t ev0, ev1;
  // FB_SPLIT: the batch is stepped as `split` env ranges ("chains") on their own streams, each chain started a few kernels
  // after the previous one, so that different stage kernels share the SMs and one chain's stragglers overlap the other's work
  int split, chain_stagger, chain_count, chain_idx; cudaStream_t cur_stream, aux[3]; cudaEvent_t chain_ev[4], join_ev[4];
  int overlap_vel; cudaEvent_t fork_ev, vel_ev;       // FB_OVERLAP_VEL: velocity stage as a parallel branch of the step (launch_step1)
  struct StepGraph { cudaGraphExec_t exec; bool seen, failed; int n_sub; long long launches; DevData d; DevModel m; };
  StepGraph graph[2];          // [hold pending?]
  bool graphs_on;
#endif
};

#ifndef FB_EMU
// FB_SPLIT chains (step_sequence): after the `chain_stagger`-th step kernel of a chain an event lets the next chain start
static void chain_mark(FbSim* s) {
  if (s->chain_count < 0) return;
  if (++s->chain_count == s->chain_stagger) cudaEventRecord(s->chain_ev[s->chain_idx], s->cur_stream);
}
#endif
static size_t slice_bytes(size_t fixed, size_t dyn_floats) { return ((((fixed + 15) & ~(size_t)15) + dyn_floats * sizeof(float)) + 15) & ~(size_t)15; }
// kernel = sequence of stages for one env.  Ph<f>: per-lane phase f(m, d, sh, e, 0, y), a warp barrier follows;
// Wf<f>: warp function f(m, d, sh, e) written with WPAR sections (its own barriers inside).
template <auto F> struct Ph {
#ifdef __CUDACC__
  template <typename Sh> static __device__ __forceinline__ void run(const DevModel& m, const DevData& d, Sh& sh, int e, int y) { F(m, d, sh, e, 0, y); }
#endif
  template <typename Sh> static void emu(const DevModel& m, const DevData& d, Sh& sh, int e) { for (int y = 0; y < FB_NY; y++) F(m, d, sh, e, 0, y); }
};
template <auto F> struct Wf {
#ifdef __CUDACC__
  template <typename Sh> static __device__ __forceinline__ void run(const DevModel& m, const DevData& d, Sh& sh, int e, int) { F(m, d, sh, e); }
#endif
  template <typename Sh> static void emu(const DevModel& m, const DevData& d, Sh& sh, int e) { F(m, d, sh, e); }
};
#ifndef FB_EMU
// one warp per env: threadIdx.x = lane ("y" of the phase functions), threadIdx.y = env within the block
template <typename Sh> __device__ __forceinline__ void set_prog(Sh&, const unsigned*) {}
__device__ __forceinline__ void set_prog(ShTree& sh, const unsigned* p) { sh.prog = p; }
template <typename Sh> __device__ __forceinline__ void set_slice(Sh&, int) {}
__device__ __forceinline__ void set_slice(ShCol& sh, int slice) { sh.slice = slice; }
template <typename Sh, typename... St>
__global__ void __launch_bounds__(32 * FB_WPB, FB_MINB) fb_run(DevModel m, DevData d, int slice, int e0, int nwarps, int blob_words) {
  extern __shared__ __align__(16) unsigned char fb_smem[];
  // optional CTA-wide copy of the sweep program in front of the warps' slices (the only block-level barrier of the step)
  unsigned* blob = reinterpret_cast<unsigned*>(fb_smem);
  if (blob_words) {
    for (int i = threadIdx.y * 32 + threadIdx.x; i < blob_words; i += 32 * FB_WPB) blob[i] = m.tsolve_blob[i];
    __syncthreads();
  }
  int e = e0 + blockIdx.x * FB_WPB + threadIdx.y;
  if (e >= e0 + nwarps) return;
  Sh& sh = *reinterpret_cast<Sh*>(fb_smem + (size_t)blob_words * 4 + (size_t)threadIdx.y * slice);
  int y = threadIdx.x;
  set_slice(sh, slice);
  if (blob_words) { set_prog(sh, blob); __syncwarp(); } else { set_prog(sh, m.tsolve_blob); __syncwarp(); }   // no CTA copy: read the program in place
#ifdef FB_CLK
  // latency profile: stage boundaries of env 0's warp, 32 slots per launch (slot 0 = kernel entry)
  int ci = 0; long long* ck = d.clk + 32 * d.clk_launch;
  if (e == 0 && y == 0) ck[ci] = clock64();
  ((St::run(m, d, sh, e, y), __syncwarp(), (e == 0 && y == 0 ? (void)(ck[++ci] = clock64()) : (void)0)), ...);
#else
  ((St::run(m, d, sh, e, y), __syncwarp()), ...);
#endif
}
template <typename Sh, typename... St>
static void fb_launch(FbSim* s, int kind, size_t dyn_floats = 0, int nwarps = -1, int blob_words = 0) {
  // step kernels (nwarps < 0) cover the env range of the chain being issued (all envs unless FB_SPLIT, see step_sequence)
  int e0 = 0;
  if (nwarps < 0) { e0 = s->cur_e0; nwarps = s->cur_n; }
#ifdef FB_CLK
  s->d.clk_launch = (int)(s->launches % 4096);
#endif
  dim3 block(32, FB_WPB), grid((nwarps + FB_WPB - 1) / FB_WPB);
  size_t slice = slice_bytes(sizeof(Sh), dyn_floats), bytes = slice * FB_WPB + (size_t)blob_words * 4;
  static size_t configured = 0;
  if (bytes > configured) { cudaFuncSetAttribute(fb_run<Sh, St...>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); configured = bytes; }
  cudaStream_t st = s->cur_stream;
  if (s->prof_on) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
    fb_run<Sh, St...><<<grid, block, bytes, st>>>(s->m, s->d, (int)slice, e0, nwarps, blob_words);
    cudaEventRecord(b, st);
    s->prof_events.push_back({kind, a, b});
  } else {
    fb_run<Sh, St...><<<grid, block, bytes, st>>>(s->m, s->d, (int)slice, e0, nwarps, blob_words);
  }
  s->launches++;
  chain_mark(s);
}
__global__ void __launch_bounds__(32 * FB_SOLVE_WPB, FB_MINB) fb_run_solve(DevModel m, DevData d, int e0, int nwarps) {
  extern __shared__ __align__(16) float fb_smem_w[];
  int e = e0 + blockIdx.x * FB_SOLVE_WPB + threadIdx.y;
  if (e >= e0 + nwarps) return;
#ifdef FB_CLK
  if (e == 0 && threadIdx.x == 0) d.clk[32 * d.clk_launch] = clock64();
#endif
  ksolve_warp(m, d, fb_smem_w + (size_t)threadIdx.y * FB_SOLVE_WARP_FLOATS, e);
#ifdef FB_CLK
  if (e == 0 && threadIdx.x == 0) d.clk[32 * d.clk_launch + 1] = clock64();
#endif
}
// heavy envs (nefc > 32) queued by fb_run_solve: one warp per env, the whole problem in shared memory; the blocks stride over the queue
__global__ void __launch_bounds__(32, 1) fb_run_solve_big(DevModel m, DevData d) {
  extern __shared__ __align__(16) float fb_smem_big[];
  const int count = *d.heavy_count;
  for (int i = blockIdx.x; i < count; i += gridDim.x) { ksolve_big(m, d, fb_smem_big, d.heavy_list[i]); __syncwarp(); }
}
static void fb_launch_warp(FbSim* s, int kind) {
#ifdef FB_CLK
  s->d.clk_launch = (int)(s->launches % 4096);
#endif
  dim3 block(32, FB_SOLVE_WPB), grid((s->cur_n + FB_SOLVE_WPB - 1) / FB_SOLVE_WPB);
  size_t bytes = sizeof(float) * FB_SOLVE_WARP_FLOATS * FB_SOLVE_WPB;
  static bool configured = false;
  if (!configured) { cudaFuncSetAttribute(fb_run_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    cudaFuncSetAttribute(fb_run_solve_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * FB_SOLVE_BIG_FLOATS)); configured = true; }
  cudaStream_t st = s->cur_stream;
  if (s->prof_on) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, st);
    fb_run_solve<<<grid, block, bytes, st>>>(s->m, s->d, s->cur_e0, s->cur_n);
    if (s->d.heavy_list) fb_run_solve_big<<<s->n_sm, 32, sizeof(float) * FB_SOLVE_BIG_FLOATS, st>>>(s->m, s->d);
    cudaEventRecord(b, st);
    s->prof_events.push_back({kind, a, b});
  } else {
    fb_run_solve<<<grid, block, bytes, st>>>(s->m, s->d, s->cur_e0, s->cur_n);
    if (s->d.heavy_list) fb_run_solve_big<<<s->n_sm, 32, sizeof(float) * FB_SOLVE_BIG_FLOATS, st>>>(s->m, s->d);
  }
  s->launches += s->d.heavy_list ? 2 : 1;
  chain_mark(s);
}
#else
template <typename Sh> static void set_prog(Sh&, const unsigned*) {}
static void set_prog(ShTree& sh, const unsigned* p) { sh.prog = p; }
template <typename Sh, typename... St>
static void fb_launch(FbSim* s, int kind, size_t dyn_floats = 0, int nwarps = -1, int blob_words = 0) {
  (void)kind;
  int e0 = 0;
  if (nwarps < 0) { e0 = s->cur_e0; nwarps = s->cur_n; }
  // the slice is exactly as large as the GPU launch makes it, followed by a canary: a phase that writes past its shared-memory
  // slice (on the GPU: into the next warp's slice or out of the block's allocation) aborts the CPU tests instead of passing
  const size_t used = slice_bytes(sizeof(Sh), dyn_floats);
  std::vector<unsigned char> buf(used + 256, 0);
  memset(buf.data() + used, 0xAB, 256);
  Sh& sh = *reinterpret_cast<Sh*>(buf.data());
  (void)blob_words; set_prog(sh, s->m.tsolve_blob);        // host emulation: the program is read in place
  for (int e = e0; e < e0 + nwarps; e++) (St::emu(s->m, s->d, sh, e), ...);
  for (size_t i = 0; i < 256; i++) if (buf[used + i] != 0xAB) { fprintf(stderr, "fb_emu: shared-memory slice overrun in a kernel of kind %d (slice %zu bytes)\n", kind, used); abort(); }
  s->launches++;
}
static void fb_launch_warp(FbSim* s, int kind) {
  (void)kind;
  static std::vector<float> buf(FB_SOLVE_WARP_FLOATS);
  for (int e = s->cur_e0; e < s->cur_e0 + s->cur_n; e++) ksolve_warp(s->m, s->d, buf.data(), e);
  s->launches++;
}
#endif

// lane == env kernels wrapped as single-phase functions
FB_DEV void ph_reset_scatter(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { kreset_scatter(m, d, e, y); }
FB_DEV void ph_scatter(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { kscatter(m, d, e, y); }
FB_DEV void ph_clear_hold(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { kclear_hold(m, d, e, y); }
// heightfield contacts: DevHf travels in its own kernel (not in DevModel / DevData: the step kernels of models without a
// heightfield keep their parameter layout)
#ifndef FB_EMU
__global__ void __launch_bounds__(32 * FB_WPB, FB_MINB) fb_hf_kernel(DevModel m, DevData d, DevHf p, int nwarps) {
  const int e = blockIdx.x * FB_WPB + threadIdx.y;
  if (e >= nwarps) return;
  khf_narrow(m, d, p, e, threadIdx.x); __syncwarp();
  khf_append(m, d, p, e, threadIdx.x);
}
#endif
static void launch_hfield(FbSim* s) {
#ifndef FB_EMU
  dim3 block(32, FB_WPB), grid((s->d.N + FB_WPB - 1) / FB_WPB);
  fb_hf_kernel<<<grid, block, 0, s->cur_stream>>>(s->m, s->d, s->hf, s->d.N);
#else
  for (int e = 0; e < s->d.N; e++) { for (int y = 0; y < FB_NY; y++) khf_narrow(s->m, s->d, s->hf, e, y); for (int y = 0; y < FB_NY; y++) khf_append(s->m, s->d, s->hf, e, y); }
#endif
  s->launches++;
}
FB_DEV void ph_task_reset(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { ktask_reset(m, d, e, y); }
FB_DEV void ph_task_reset2(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { ktask_reset2(m, d, e, y); }
FB_DEV void ph_task_before(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { ktask_before(m, d, e, y); }
FB_DEV void ph_task_commit(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { ktask_commit(m, d, e, y); }
FB_DEV void ph_task_after(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { ktask_after(m, d, e, y); }
FB_DEV void ph_pack(const DevModel& m, const DevData& d, ShNone&, int e, int, int y) { kpack(m, d, e, y, d.nsub_done > 0 ? 1.0f / d.nsub_done : 1.0f); ktaskobs(m, d, e, y); }
// Smooth dynamics in half-solve form: with M = L^T D L and u = L^-T qfrc_smooth,
//   J qacc_smooth = Z (D^-1/2 u)          (Z = D^-1/2 L^-T J^T, rows written by the projection kernel)
//   qacc          = L^-1 (D^-1 u + D^-1/2 Z^T f)
// so the step needs L^-T once (here) and L^-1 once (finish kernel) instead of two full solves.  qtmp <- D^-1 u,
// dof_isd <- D^-1/2, XS <- D^-1/2 u for kref.
FB_WARPFN void wf_smooth_solve(const DevModel& m, const DevData& d, ShTree& sh, int e) { tsolve_a(m, d, sh, e); }
FB_DEV void ph_smooth_out(FB_PHASE_ARGS) {
  float* xs = sh_dyn(sh); const float* lds = xs + FB_NXS(m);
  for (int i = y; i < m.nv; i += FB_NY) {
    float D = LDS(m.dof_Madr[i]), u = XS(i), isd = 1.0f / sqrtf(D);
    AT(d.qtmp, i) = u / D; AT(d.dof_isd, i) = isd; XS(i) = u * isd;
  }
}

// the stage lists of the seven step kernels (shared by the one-kernel-per-stage launches and the fused launches below)
#define FB_ST_POS Ph<kpos_p0>, Ph<kpos_p1>, Ph<kpos_p1b>, Ph<kpos_p2>, Ph<kpos_p3>, Ph<kpos_p3b>, Ph<kpos_p4>, Wf<kpos_factor>, Ph<kpos_p6w>, Ph<kpos_p6d>, Wf<kpos_factor>, Ph<kpos_p9>
#define FB_ST_COL Ph<kcol_stage>, Wf<kcol_broad>, Ph<kcol_narrow>, Wf<kcol_mpr>, Ph<kcol_compact>
#define FB_ST_PROJ Ph<kcon_p0>, Ph<kcon_p1>, Ph<kcon_p2>, Ph<kcon_p3>, Ph<kproj_p0>, Ph<kproj_p1>
#define FB_ST_VEL Ph<kvel_p0>, Ph<kvel_p1>, Ph<kvel_p1b>, Ph<kvel_p2>, Ph<kvel_p3>, Ph<kvel_p3b>, Ph<kvel_p4>
#define FB_ST_SMOOTH Ph<kact_p0>, Ph<kact_p1>, Ph<kact_p2>, Ph<kact_p2b>, Ph<kact_p3>, Wf<wf_smooth_solve>, Ph<ph_smooth_out>, Ph<kref>
#define FB_ST_FINISH Ph<kfin_f1>, Wf<kfin_solve>, Ph<kfin_f5>, Ph<kfin_f6>, Ph<kfin_f7>, Wf<kfin_solve_euler>, Ph<kfin_f8>, Ph<kfin_f9>, Ph<kfin_f10>
static size_t dyn_pos(const DevModel& m) { return (size_t)FB_PARTF + std::max((size_t)m.nM, (size_t)10 * m.nbody); }   // LS region: inertia matrix rows, before that the composite inertias (CRBS)
static size_t dyn_col(const DevModel& m) { return (size_t)FB_COL_DYN(m); }
static size_t dyn_proj(const DevModel&) { return (size_t)FB_NY * FB_ZCAP; }
static size_t dyn_vel(const DevModel&) { return (size_t)FB_PARTF; }
static size_t dyn_tsolve(const DevModel& m) { return (size_t)(FB_NXS(m) + ((m.nM + 3) & ~3)); }

static void launch_step1(FbSim* s) {
  fb_launch<ShTree, FB_ST_POS>(s, K_POS, dyn_pos(s->m));
#ifndef FB_EMU
  // The velocity stage (comVel, RNE bias, passive forces, velocity sensors) needs the position stage only; collision and the
  // constraint rows need it too but not each other's outputs: vel runs as a parallel branch next to col -> proj (second stream,
  // captured into the step graph as a fork / join).  Every kernel is a single wave that fills the register file, so the branch's
  // blocks start as the collision kernel's blocks drain -- they fill its barrier / MPR tail instead of waiting behind proj.
  if (s->overlap_vel && s->split == 1 && !s->prof_on) {
    cudaStream_t main_st = s->cur_stream, side = s->aux[2];
    cudaEventRecord(s->fork_ev, main_st); cudaStreamWaitEvent(side, s->fork_ev, 0);
    s->cur_stream = side;
    fb_launch<ShTree, FB_ST_VEL>(s, K_VEL, dyn_vel(s->m));
    cudaEventRecord(s->vel_ev, side);
    s->cur_stream = main_st;
    fb_launch<ShCol, FB_ST_COL>(s, K_COL, dyn_col(s->m));
    if (s->hf_on) launch_hfield(s);
    fb_launch<ShCon, FB_ST_PROJ>(s, K_PROJ, dyn_proj(s->m));
    cudaStreamWaitEvent(main_st, s->vel_ev, 0);
    return;
  }
#endif
  fb_launch<ShCol, FB_ST_COL>(s, K_COL, dyn_col(s->m));
  if (s->hf_on) launch_hfield(s);              // heightfield contacts join the contact list before the constraint rows are built
  fb_launch<ShCon, FB_ST_PROJ>(s, K_PROJ, dyn_proj(s->m));
  fb_launch<ShTree, FB_ST_VEL>(s, K_VEL, dyn_vel(s->m));
}
static void launch_step2(FbSim* s, bool integrate) {
  s->d.do_integrate = integrate ? 1 : 0;     // read by the solve (warm-start bookkeeping) and the finish kernel
  fb_launch<ShTree, FB_ST_SMOOTH>(s, K_SMOOTH, dyn_tsolve(s->m), -1, s->blob_in_smem ? s->m.ts_blob_words : 0);
  fb_launch_warp(s, K_SOLVE);
  fb_launch<ShTree, FB_ST_FINISH>(s, K_FINISH, dyn_tsolve(s->m), -1, s->blob_in_smem ? s->m.ts_blob_words : 0);
}

// -------------------------------------------------------------------------------------------
// Fused launches.  An env is owned by one warp from the first to the last stage of a step and no stage looks at another
// env, so consecutive stage kernels can run back to back inside one launch: a *group* is one of the kernels above (its
// shared-memory struct re-interpreted on the warp's slice, sized for the largest group), a fused kernel is a list of groups,
// optionally looped over the substeps of a control step.  What it buys: no grid-wide drain between stages (the envs of a
// batch differ in contact count and Newton iterations, so every stage boundary otherwise waits for its slowest warp), the
// warps of an SM drift apart and stop competing for the same unit at the same time, and an env's intermediates are re-read
// from the L1/L2 of the SM that wrote them.  FB_FUSE selects: 0 one kernel per stage (7 per substep), 1 two kernels per
// substep ([smooth solve finish] [pos col proj vel]), 2 one kernel per substep, 3 one kernel per control step.
template <typename Sh, typename... St> struct Grp {
#ifdef __CUDACC__
  static __device__ __forceinline__ void run(const DevModel& m, const DevData& d, unsigned char* slice, int stride, const unsigned* prog, int e, int y) {
    Sh& sh = *reinterpret_cast<Sh*>(slice);
    set_slice(sh, stride); set_prog(sh, prog); __syncwarp();
    ((St::run(m, d, sh, e, y), __syncwarp()), ...);
  }
#endif
  static void emu(const DevModel& m, const DevData& d, unsigned char* slice, int e) {
    Sh& sh = *reinterpret_cast<Sh*>(slice);
    set_prog(sh, m.tsolve_blob);
    (St::emu(m, d, sh, e), ...);
  }
};
struct GrpSolve {
#ifdef __CUDACC__
  static __device__ __forceinline__ void run(const DevModel& m, const DevData& d, unsigned char* slice, int, const unsigned*, int e, int) {
    ksolve_warp(m, d, reinterpret_cast<float*>(slice), e); __syncwarp();
  }
#else
  static void emu(const DevModel& m, const DevData& d, unsigned char* slice, int e) { ksolve_warp(m, d, reinterpret_cast<float*>(slice), e); }
#endif
};
using GSmooth = Grp<ShTree, FB_ST_SMOOTH>; using GFinish = Grp<ShTree, FB_ST_FINISH>; using GPos = Grp<ShTree, FB_ST_POS>;
using GCol = Grp<ShCol, FB_ST_COL>; using GProj = Grp<ShCon, FB_ST_PROJ>; using GVel = Grp<ShTree, FB_ST_VEL>;
FB_DEV void fused_sens_accum(const DevModel& m, const DevData& d, int e, int y, int first) {
  for (int i = y; i < m.nsensordata; i += FB_NY) AT(d.sensor_sum, i) = (first ? 0.0f : AT(d.sensor_sum, i)) + AT(d.sensordata, i);
}
static size_t fused_slice_bytes(const DevModel& m) {
  size_t b = slice_bytes(sizeof(ShTree), std::max(std::max(dyn_pos(m), dyn_tsolve(m)), dyn_vel(m)));
  b = std::max(b, slice_bytes(sizeof(ShCol), dyn_col(m)));
  b = std::max(b, slice_bytes(sizeof(ShCon), dyn_proj(m)));
  b = std::max(b, slice_bytes(0, (size_t)FB_SOLVE_WARP_FLOATS));
  return b;
}
#ifndef FB_EMU
// loop_sens: the kernel runs n_sub whole substeps and restarts / continues the sensor sums itself (d.sens_mode is -1 then)
template <typename... G>
__global__ void __launch_bounds__(32 * FB_WPB, FB_MINB) fb_run_fused(DevModel m, DevData d, int slice, int nwarps, int blob_words, int n_sub, int loop_sens) {
  extern __shared__ __align__(16) unsigned char fb_smem[];
  unsigned* blob = reinterpret_cast<unsigned*>(fb_smem);
  if (blob_words) {
    for (int i = threadIdx.y * 32 + threadIdx.x; i < blob_words; i += 32 * FB_WPB) blob[i] = m.tsolve_blob[i];
    __syncthreads();
  }
  int e = blockIdx.x * FB_WPB + threadIdx.y;
  if (e >= nwarps) return;
  unsigned char* sl = fb_smem + (size_t)blob_words * 4 + (size_t)threadIdx.y * slice;
  const unsigned* prog = blob_words ? blob : m.tsolve_blob;
  const int y = threadIdx.x;
  for (int k = 0; k < n_sub; k++) {
    (G::run(m, d, sl, slice, prog, e, y), ...);
    if (loop_sens) { fused_sens_accum(m, d, e, y, k == 0); __syncwarp(); }
  }
}
#endif
template <typename... G>
static void fb_launch_fused(FbSim* s, int kind, int n_sub, int loop_sens) {
  const int nwarps = s->d.Np;
  size_t slice = fused_slice_bytes(s->m);
#ifndef FB_EMU
  const int blob_words = s->blob_in_smem ? s->m.ts_blob_words : 0;
  dim3 block(32, FB_WPB), grid((nwarps + FB_WPB - 1) / FB_WPB);
  size_t bytes = slice * FB_WPB + (size_t)blob_words * 4;
  static size_t configured = 0;
  if (bytes > configured) { cudaFuncSetAttribute(fb_run_fused<G...>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); configured = bytes; }
  cudaEvent_t a = nullptr, b = nullptr;
  if (s->prof_on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s->cur_stream); }
  fb_run_fused<G...><<<grid, block, bytes, s->cur_stream>>>(s->m, s->d, (int)slice, nwarps, blob_words, n_sub, loop_sens);
  if (s->prof_on) { cudaEventRecord(b, s->cur_stream); s->prof_events.push_back({kind, a, b}); }
#else
  (void)kind;
  static std::vector<unsigned char> buf;
  if (buf.size() < slice + 64) buf.resize(slice + 64);
  for (int e = 0; e < nwarps; e++)
    for (int k = 0; k < n_sub; k++) {
      (G::emu(s->m, s->d, buf.data(), e), ...);
      if (loop_sens) for (int y = 0; y < FB_NY; y++) fused_sens_accum(s->m, s->d, e, y, k == 0);
    }
#endif
  s->launches++;
}

// -------------------------------------------------------------------------------------------
// model upload
template <typename T> static const T* up(FbSim* s, const std::vector<T>& v) {
  void* p = dev_alloc(sizeof(T) * std::max<size_t>(v.size(), 1));
  if (v.size()) h2d(p, v.data(), sizeof(T) * v.size());
  s->allocs.push_back(p);
  return (const T*)p;
}
static const float* upf(FbSim* s, const double* src, size_t n) { std::vector<float> v(n); for (size_t i = 0; i < n; i++) v[i] = (float)src[i]; return up(s, v); }
static const float* upf3(FbSim* s, const double* src, size_t n) {      // n 3-vectors, padded to four floats each (fb_math.h: mld3)
  std::vector<float> v(4 * std::max<size_t>(n, 1), 0.0f);
  for (size_t i = 0; i < n; i++) for (int k = 0; k < 3; k++) v[4 * i + k] = (float)src[3 * i + k];
  return up(s, v);
}
static const int* upi(FbSim* s, const int32_t* src, size_t n) { std::vector<int> v(src, src + n); return up(s, v); }
template <typename T> static T* dalloc(FbSim* s, size_t n) { void* p = dev_alloc(sizeof(T) * n); s->allocs.push_back(p); return (T*)p; }
// free the buffers a replaced program owned (the stream is idle: the callers synchronise first)
static void scope_free(FbSim* s, std::vector<void*>& scope) {
  for (void* p : scope) { if (!p) continue; for (auto& a : s->allocs) if (a == p) { a = nullptr; break; } dev_free(p); }
  scope.clear();
}

static int build_model(FbSim* s, const FbModel* h) {
  DevModel& m = s->m;
  memset(&m, 0, sizeof(m));
  m.nq = h->nq; m.nv = h->nv; m.nu = h->nu; m.na = h->na; m.nbody = h->nbody; m.njnt = h->njnt; m.ngeom = h->ngeom;
  m.npair = h->npair; m.nsite = h->nsite; m.ntendon = h->ntendon; m.nwrap = h->nwrap; m.nsensor = h->nsensor;
  m.nsensordata = h->nsensordata; m.nM = h->nM; m.nfluid = h->nfluid;
  m.noslip_iterations = h->opt_noslip_iterations; m.cone_elliptic = h->opt_cone_elliptic; m.ls_tolerance = 1e-3f;
  // solver limits = the model's own (mjOption iterations / ls_iterations: 100 / 50 for the fly, fruitfly.xml:4 leaves the defaults);
  // the loops leave on convergence long before (1-4 Newton iterations, 2-4 line-search evaluations per substep in the walking workload)
  m.max_iter = h->opt_iterations > 0 ? h->opt_iterations : 100; m.ls_iter = h->opt_ls_iterations > 0 ? h->opt_ls_iterations : 50;
  if (h->nv > 4 * FB_SOLVE_NCAP) { s->err = "nv exceeds the solver's per-dof accumulator window (4 * FB_SOLVE_NCAP)"; return -3; }
  { const char* rt = getenv("FB_SOLVE_RTOL"); m.solve_rtol = rt ? (float)atof(rt) : 1e-6f; }     // relative improvement that ends the Newton iteration (test hook)
  { const char* nc = getenv("FB_SOLVE_NCAP"); m.solve_ncap = nc ? atoi(nc) : FB_SOLVE_NCAP; if (m.solve_ncap > FB_SOLVE_NCAP) m.solve_ncap = FB_SOLVE_NCAP; }   // test hook: smaller cap -> global-memory path
  m.timestep = (float)h->opt_timestep; m.density = (float)h->opt_density; m.viscosity = (float)h->opt_viscosity;
  for (int i = 0; i < 3; i++) { m.gravity[i] = (float)h->opt_gravity[i]; m.wind[i] = (float)h->opt_wind[i]; }
  m.impratio = (float)h->opt_impratio; m.tolerance = (float)h->opt_tolerance; m.noslip_tolerance = (float)h->opt_noslip_tolerance;
  m.meaninertia = (float)h->stat_meaninertia;
  int nb = m.nbody, nv = m.nv;
  // ---- tree partition: roots = children of world; each root's child sub-trees are packed into lists
  std::vector<int> roots, isroot(nb, 0);
  for (int b = 1; b < nb; b++) if (h->body_parentid[b] == 0) { roots.push_back(b); isroot[b] = 1; }
  isroot[0] = 1;   // the world behaves like a root for "parent is root" tests (never accumulated into)
  for (int r : roots) {
    if (!(h->body_jntnum[r] == 0 || (h->body_jntnum[r] == 1 && h->jnt_type[h->body_jntadr[r]] == FB_JNT_FREE))) { s->err = "root bodies must carry a single free joint or none"; return -3; }
  }
  std::vector<std::vector<int>> subtrees; std::vector<int> sub_root;
  for (size_t ri = 0; ri < roots.size(); ri++) {
    for (int b = 1; b < nb; b++) if (h->body_parentid[b] == roots[ri]) {
      std::vector<int> st; std::vector<char> in(nb, 0); in[b] = 1; st.push_back(b);
      for (int c = b + 1; c < nb; c++) if (in[h->body_parentid[c]]) { in[c] = 1; st.push_back(c); }
      subtrees.push_back(st); sub_root.push_back((int)ri);
    }
  }
  size_t maxsz = 1, total = 0; for (auto& st : subtrees) { maxsz = std::max(maxsz, st.size()); total += st.size(); }
  int nlist = (int)std::min<size_t>(FB_NLMAX, std::max<size_t>(1, (total + maxsz - 1) / maxsz + 1));
  std::vector<std::vector<int>> lists(nlist); std::vector<int> lroot(nlist, -1);
  std::vector<size_t> order(subtrees.size()); for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return subtrees[a].size() > subtrees[b].size(); });
  for (size_t oi : order) {
    int best = -1;
    for (int l = 0; l < nlist; l++) { if (lroot[l] != -1 && lroot[l] != sub_root[oi]) continue; if (best < 0 || lists[l].size() < lists[best].size()) best = l; }
    if (best < 0) { s->err = "tree partition: more roots with children than lists"; return -3; }
    lroot[best] = sub_root[oi];
    lists[best].insert(lists[best].end(), subtrees[oi].begin(), subtrees[oi].end());
  }
  std::vector<int> ladr, lnum, lbody, lr;
  for (int l = 0; l < nlist; l++) { ladr.push_back((int)lbody.size()); lnum.push_back((int)lists[l].size()); lbody.insert(lbody.end(), lists[l].begin(), lists[l].end()); lr.push_back(lroot[l] < 0 ? 0 : lroot[l]); }
  m.nroot = (int)roots.size(); m.nlist = nlist;
  { // factorisation order of each list: bodies in reverse, dofs in reverse
    std::vector<int> dadr, dnum, dl; int mx = 0;
    for (int l = 0; l < nlist; l++) { dadr.push_back((int)dl.size()); int c = 0;
      for (int bi = (int)lists[l].size() - 1; bi >= 0; bi--) { int bb = lists[l][bi]; for (int kk = h->body_dofnum[bb] - 1; kk >= 0; kk--) { dl.push_back(h->body_dofadr[bb] + kk); c++; } }
      dnum.push_back(c); mx = std::max(mx, c); }
    m.list_dofadr = up(s, dadr); m.list_ndof = up(s, dnum); m.list_dof = up(s, dl); m.max_list_ndof = mx;
  }
  m.root_body = up(s, roots); m.list_adr = up(s, ladr); m.list_num = up(s, lnum); m.list_body = up(s, lbody); m.list_root = up(s, lr);
  m.body_isroot = up(s, isroot);
  { // bodies the acceleration-stage sensors need: force sensors read the subtree sum at the site's body;
    // accelerometers need the (delta) acceleration of the site's body; both need their ancestors' deltas
    std::vector<int> sfrc(nb, 0), sacc(nb, 0), frcroot(nb, 0);
    for (int t = 0; t < h->nsensor; t++) { int b = h->site_bodyid[h->sensor_objid[t]]; if (h->sensor_type[t] == FB_SENS_FORCE) frcroot[b] = 1; if (h->sensor_type[t] == FB_SENS_FORCE || h->sensor_type[t] == FB_SENS_ACCELEROMETER) sacc[b] = 1; }
    for (int b = 1; b < nb; b++) if (frcroot[b] || sfrc[h->body_parentid[b]]) sfrc[b] = 1;
    for (int b = 1; b < nb; b++) if (sfrc[b]) sacc[b] = 1;
    for (int b = nb - 1; b > 0; b--) if (sacc[b]) sacc[h->body_parentid[b]] = 1;
    sacc[0] = 0;
    m.body_sensacc = up(s, sacc); m.body_sensfrc = up(s, sfrc);
  }
  // geoms / sites per body (both are stored in body order by the compiler)
  std::vector<int> gadr(nb, 0), gnum(nb, 0), sadr(nb, 0), snum(nb, 0);
  for (int g = 0; g < m.ngeom; g++) { int b = h->geom_bodyid[g]; if (gnum[b] == 0) gadr[b] = g; gnum[b]++; if (g > 0 && h->geom_bodyid[g] < h->geom_bodyid[g - 1]) { s->err = "geoms not in body order"; return -3; } }
  for (int t = 0; t < m.nsite; t++) { int b = h->site_bodyid[t]; if (snum[b] == 0) sadr[b] = t; snum[b]++; if (t > 0 && h->site_bodyid[t] < h->site_bodyid[t - 1]) { s->err = "sites not in body order"; return -3; } }
  m.body_geomadr = up(s, gadr); m.body_geomnum = up(s, gnum); m.body_siteadr = up(s, sadr); m.body_sitenum = up(s, snum);
  // dof helpers
  std::vector<int> subend(nv), depth(nv), disroot(nv), chainlen(nv);
  for (int i = 0; i < nv; i++) subend[i] = i;
  for (int i = nv - 1; i >= 0; i--) { int p = h->dof_parentid[i]; if (p >= 0) subend[p] = std::max(subend[p], subend[i]); }
  for (int i = 0; i < nv; i++) {
    disroot[i] = isroot[h->dof_bodyid[i]] && h->dof_bodyid[i] != 0;
    int len = 0, nonroot = 0;
    for (int j = i; j >= 0; j = h->dof_parentid[j]) { len++; if (j != i && !(isroot[h->dof_bodyid[j]] && h->dof_bodyid[j] != 0)) nonroot++; }
    chainlen[i] = len;
    depth[i] = disroot[i] ? (i - h->body_dofadr[h->dof_bodyid[i]]) : nonroot;
  }
  { std::vector<int> anc(h->nM, -1); for (int i = 0; i < nv; i++) { int t = 0; for (int j = i; j >= 0; j = h->dof_parentid[j], t++) anc[h->dof_Madr[i] + t] = j; } m.dof_anc = up(s, anc); }
  { // shared-memory slot of the t-th ancestor for the triangular solves: root dofs map to the list's private accumulators
    std::vector<int> dof_list(nv, -1);
    for (int l = 0; l < nlist; l++) for (int bb : lists[l]) for (int kk = 0; kk < h->body_dofnum[bb]; kk++) dof_list[h->body_dofadr[bb] + kk] = l;
    std::vector<int> slot(h->nM, 0);
    for (int i = 0; i < nv; i++) { int t = 0; for (int j = i; j >= 0; j = h->dof_parentid[j], t++) slot[h->dof_Madr[i] + t] = (disroot[j] && dof_list[i] >= 0) ? nv + FB_ROOTD * dof_list[i] + depth[j] : j; }
    m.dof_ancslot = up(s, slot);
    std::vector<int> rootidx(nv, -1), haslists(roots.size(), 0);
    for (int i = 0; i < nv; i++) if (dof_list[i] >= 0) { rootidx[i] = lr[dof_list[i]]; haslists[lr[dof_list[i]]] = 1; }
    m.dof_rootidx = up(s, rootidx); m.root_haslists = up(s, haslists);
  }
  { // common-ancestor counts of every dof pair (kproj_p1: the dofs two constraint rows share are the common TAIL of their chains)
    if (nv > 255) { s->err = "too many dofs for the byte-sized common-ancestor table"; return -3; }
    if (nv > 128) { s->err = "too many dofs for the register-path force gather (4 dofs per lane, fb_solver_reg.h)"; return -3; }
    std::vector<unsigned char> lca((size_t)nv * nv, 0); std::vector<int> mark(nv, -1);
    for (int a = 0; a < nv; a++) {
      for (int j = a; j >= 0; j = h->dof_parentid[j]) mark[j] = a;
      for (int b = 0; b < nv; b++) { int j = b; while (j >= 0 && mark[j] != a) j = h->dof_parentid[j]; lca[(size_t)a * nv + b] = (unsigned char)(j >= 0 ? chainlen[j] : 0); }
    }
    m.dof_lca = up(s, lca);
  }
  { std::vector<int> adh(nb, -1); for (int i = 0; i < h->nu; i++) if (h->actuator_trntype[i] == FB_TRN_BODY) adh[h->actuator_trnid[i]] = i; m.body_adhesion = up(s, adh); }
  { // packed headers of the lock-step sweeps (see fb_tree.h) and the row address of every ancestor entry
    if (h->nM >= 4096 || nv >= 256) { s->err = "model too large for the packed sweep headers (nM < 4096, nv < 256)"; return -3; }
    std::vector<unsigned> ha((size_t)FB_NY * std::max(m.max_list_ndof, 1), 0xffffffffu), hc(ha);
    std::vector<int> dadr_h, dnum_h, dl_h;      // rebuild the per-list dof order exactly as above
    for (int l = 0; l < nlist; l++) { dadr_h.push_back((int)dl_h.size()); int c = 0;
      for (int bi = (int)lists[l].size() - 1; bi >= 0; bi--) { int bb = lists[l][bi]; for (int kk = h->body_dofnum[bb] - 1; kk >= 0; kk--) { dl_h.push_back(h->body_dofadr[bb] + kk); c++; } }
      dnum_h.push_back(c); }
    auto pack = [&](int k) -> unsigned {
      if (chainlen[k] >= 64 || depth[k] >= 64) return 0xfffffffeu;
      return (unsigned)h->dof_Madr[k] | ((unsigned)chainlen[k] << 12) | ((unsigned)depth[k] << 18) | ((unsigned)k << 24); };
    for (int l = 0; l < nlist && FB_FSUB * l + FB_FSUB <= FB_NY; l++) for (int st = 0; st < dnum_h[l]; st++) for (int u = 0; u < FB_FSUB; u++) {
      unsigned a = pack(dl_h[dadr_h[l] + st]), c = pack(dl_h[dadr_h[l] + dnum_h[l] - 1 - st]);
      if (a == 0xfffffffeu || c == 0xfffffffeu) { s->err = "chain too long for the packed sweep headers"; return -3; }
      ha[(size_t)st * FB_NY + FB_FSUB * l + u] = a; hc[(size_t)st * FB_NY + FB_FSUB * l + u] = c; }
    m.step_hdr_a = up(s, ha); m.step_hdr_c = up(s, hc);
    { // the same headers + ancestor slots as bytes, in one blob that the smooth / finish kernels copy into shared memory
      if (nv + FB_ROOTD * nlist >= 256) { s->err = "too many dofs for byte-sized sweep slots"; return -3; }
      const int H = (int)ha.size(), nmp = (h->nM + 3) & ~3;
      std::vector<unsigned> blob((((size_t)2 * H + 2 * nmp / 4) + 3) & ~(size_t)3, 0u);       // whole 16-byte units: the warps' slices follow
      for (int i = 0; i < H; i++) { blob[i] = ha[i]; blob[H + i] = hc[i]; }
      unsigned char* b8 = reinterpret_cast<unsigned char*>(blob.data() + 2 * H);
      std::vector<int> dof_list2(nv, -1);
      for (int l = 0; l < nlist; l++) for (int bb : lists[l]) for (int kk = 0; kk < h->body_dofnum[bb]; kk++) dof_list2[h->body_dofadr[bb] + kk] = l;
      for (int i = 0; i < nv; i++) { int t = 0; for (int j = i; j >= 0; j = h->dof_parentid[j], t++) {
          int slot = (disroot[j] && dof_list2[i] >= 0) ? nv + FB_ROOTD * dof_list2[i] + depth[j] : j;
          b8[h->dof_Madr[i] + t] = (unsigned char)slot; b8[nmp + h->dof_Madr[i] + t] = (unsigned char)j; } }
      m.tsolve_blob = up(s, blob); m.ts_hdr_words = H; m.ts_nm_pad = nmp; m.ts_blob_words = (int)blob.size();
    }
    std::vector<int> ancadr(h->nM, 0);
    for (int i = 0; i < nv; i++) { int t = 0; for (int j = i; j >= 0; j = h->dof_parentid[j], t++) ancadr[h->dof_Madr[i] + t] = h->dof_Madr[j]; }
    m.M_ancadr = up(s, ancadr);
    { // balanced schedule of the factorisation sweeps (fb_tree.h: factor_step_sched): the rank-1 update of one ancestor row by the
      // step's dof of one list is a work item (cost = its length + a constant); the items of a step touch different rows, so they
      // are dealt to the 32 lanes longest-first onto the least loaded lane instead of 3 fixed lanes per list
      const int nstep = std::max(m.max_list_ndof, 1);
      std::vector<unsigned> rng((size_t)FB_NY * nstep, 0u); std::vector<unsigned> items;
      for (int st = 0; st < nstep; st++) {
        struct It { unsigned x, y; int cost; };
        std::vector<It> its;
        for (int l = 0; l < nlist; l++) {
          if (st >= dnum_h[l]) continue;
          const int k = dl_h[dadr_h[l] + st], adrk = h->dof_Madr[k], len = chainlen[k];
          for (int t = 1; t < 1 + depth[k]; t++) its.push_back({(unsigned)adrk | ((unsigned)t << 12) | ((unsigned)(len - t) << 18), (unsigned)ancadr[adrk + t], len - t + 6});
        }
        std::stable_sort(its.begin(), its.end(), [](const It& a, const It& b) { return a.cost > b.cost; });
        std::vector<std::vector<It>> lane_items(FB_NY); std::vector<int> load(FB_NY, 0);
        for (const It& it : its) { int best = 0; for (int y = 1; y < FB_NY; y++) if (load[y] < load[best]) best = y; lane_items[best].push_back(it); load[best] += it.cost; }
        for (int y = 0; y < FB_NY; y++) {
          if (items.size() / 2 >= 65536 || lane_items[y].size() >= 65536) { s->err = "factorisation schedule too long"; return -3; }
          rng[(size_t)st * FB_NY + y] = (unsigned)(items.size() / 2) | ((unsigned)lane_items[y].size() << 16);
          for (const It& it : lane_items[y]) { items.push_back(it.x); items.push_back(it.y); }
        }
      }
      if (items.empty()) { items.push_back(0); items.push_back(0); }
      m.fs_rng = up(s, rng); m.fs_items = reinterpret_cast<const FsItem*>(up(s, items));
    }
  }
  { std::vector<float> mdamp(h->nM, 0.0f); for (int i = 0; i < nv; i++) mdamp[h->dof_Madr[i]] = (float)h->dof_damping[i]; m.M_damp = up(s, mdamp); }
  m.dof_subend = up(s, subend); m.dof_depth = up(s, depth); m.dof_isroot = up(s, disroot); m.dof_chainlen = up(s, chainlen);
  // plain copies
  m.body_parentid = upi(s, h->body_parentid, nb); m.body_rootid = upi(s, h->body_rootid, nb);
  m.body_jntadr = upi(s, h->body_jntadr, nb); m.body_jntnum = upi(s, h->body_jntnum, nb);
  m.body_dofadr = upi(s, h->body_dofadr, nb); m.body_dofnum = upi(s, h->body_dofnum, nb);
  m.body_lastdof = upi(s, h->body_lastdof, nb); m.body_fluid_ellipsoid = upi(s, h->body_fluid_ellipsoid, nb);
  m.body_pos = upf3(s, h->body_pos, nb); m.body_quat = upf(s, h->body_quat, 4 * nb); m.body_ipos = upf3(s, h->body_ipos, nb);
  m.body_iquat = upf(s, h->body_iquat, 4 * nb); m.body_mass = upf(s, h->body_mass, nb); m.body_inertia = upf3(s, h->body_inertia, nb);
  m.body_invweight0 = upf(s, h->body_invweight0, 2 * nb);
  int nj = m.njnt;
  m.jnt_type = upi(s, h->jnt_type, nj); m.jnt_qposadr = upi(s, h->jnt_qposadr, nj); m.jnt_dofadr = upi(s, h->jnt_dofadr, nj);
  m.jnt_bodyid = upi(s, h->jnt_bodyid, nj); m.jnt_limited = upi(s, h->jnt_limited, nj);
  m.jnt_pos = upf3(s, h->jnt_pos, nj); m.jnt_axis = upf3(s, h->jnt_axis, nj); m.jnt_stiffness = upf(s, h->jnt_stiffness, nj);
  m.jnt_range = upf(s, h->jnt_range, 2 * nj); m.jnt_solref = upf(s, h->jnt_solref, 2 * nj); m.jnt_solimp = upf(s, h->jnt_solimp, 5 * nj);
  m.jnt_margin = upf(s, h->jnt_margin, nj); m.qpos0 = upf(s, h->qpos0, m.nq); m.qpos_spring = upf(s, h->qpos_spring, m.nq);
  m.dof_bodyid = upi(s, h->dof_bodyid, nv); m.dof_jntid = upi(s, h->dof_jntid, nv); m.dof_parentid = upi(s, h->dof_parentid, nv);
  m.dof_Madr = upi(s, h->dof_Madr, nv); m.dof_armature = upf(s, h->dof_armature, nv); m.dof_damping = upf(s, h->dof_damping, nv);
  m.dof_invweight0 = upf(s, h->dof_invweight0, nv);
  int ng = m.ngeom;
  m.geom_type = upi(s, h->geom_type, ng); m.geom_bodyid = upi(s, h->geom_bodyid, ng); m.geom_condim = upi(s, h->geom_condim, ng);
  m.geom_size = upf3(s, h->geom_size, ng); m.geom_pos = upf3(s, h->geom_pos, ng); m.geom_quat = upf(s, h->geom_quat, 4 * ng);
  m.geom_rbound = upf(s, h->geom_rbound, ng); m.geom_friction = upf(s, h->geom_friction, 3 * ng); m.geom_solmix = upf(s, h->geom_solmix, ng);
  m.geom_solref = upf(s, h->geom_solref, 2 * ng); m.geom_solimp = upf(s, h->geom_solimp, 5 * ng);
  m.geom_margin = upf(s, h->geom_margin, ng); m.geom_gap = upf(s, h->geom_gap, ng);
  m.pair_geom1 = upi(s, h->pair_geom1, m.npair); m.pair_geom2 = upi(s, h->pair_geom2, m.npair);
  { // packed broadphase record per pair: geoms + plane flag, and margin + bounding radii (plane pairs: of geom 2 only)
    std::vector<int> pi((size_t)std::max(m.npair, 1), 0); std::vector<float> rs(std::max(m.npair, 1), 0.0f);
    for (int k = 0; k < m.npair; k++) {
      int g1 = h->pair_geom1[k], g2 = h->pair_geom2[k]; bool plane = h->geom_type[g1] == FB_GEOM_PLANE;
      if (g1 >= 32768 || g2 >= 32768) { s->err = "too many geoms for the packed pair record"; return -3; }
      pi[k] = g1 | (g2 << 15) | (plane ? (1 << 30) : 0);
      double margin = std::max(h->geom_margin[g1], h->geom_margin[g2]);
      rs[k] = (float)(margin + h->geom_rbound[g2] + (plane ? 0.0 : h->geom_rbound[g1]));
    }
    m.pair_info = up(s, pi); m.pair_rsum = up(s, rs);
  }
  {  // chunk boundaries: plane pairs produce most contacts (up to 4 each), weight them 16x
    m.nchunk = FB_MAXCHUNK;
    std::vector<int> w(m.npair); long tot = 0;
    for (int k = 0; k < m.npair; k++) { w[k] = (h->geom_type[h->pair_geom1[k]] == FB_GEOM_PLANE) ? 16 : 1; tot += w[k]; }
    std::vector<int> cs(m.nchunk + 1, m.npair); cs[0] = 0;
    long acc = 0; int c = 1;
    for (int k = 0; k < m.npair && c < m.nchunk; k++) { acc += w[k]; if (acc * m.nchunk >= tot * c) { cs[c++] = k + 1; } }
    for (; c <= m.nchunk; c++) cs[c] = m.npair;
    m.chunk_start = up(s, cs);

  }
  m.fluid_bodyid = upi(s, h->fluid_bodyid, m.nfluid); m.fluid_pos = upf3(s, h->fluid_pos, m.nfluid);
  m.fluid_quat = upf(s, h->fluid_quat, 4 * m.nfluid); m.fluid_size = upf3(s, h->fluid_size, m.nfluid); m.fluid_coef = upf(s, h->fluid_coef, 12 * m.nfluid);
  m.site_bodyid = upi(s, h->site_bodyid, m.nsite); m.site_type = upi(s, h->site_type, m.nsite);
  m.site_pos = upf3(s, h->site_pos, m.nsite); m.site_quat = upf(s, h->site_quat, 4 * m.nsite); m.site_size = upf3(s, h->site_size, m.nsite);
  m.tendon_adr = upi(s, h->tendon_adr, m.ntendon); m.tendon_num = upi(s, h->tendon_num, m.ntendon);
  m.wrap_dofid = upi(s, h->wrap_dofid, m.nwrap); m.wrap_qposadr = upi(s, h->wrap_qposadr, m.nwrap); m.wrap_coef = upf(s, h->wrap_coef, m.nwrap);
  int nu = m.nu;
  m.actuator_trntype = upi(s, h->actuator_trntype, nu); m.actuator_trnid = upi(s, h->actuator_trnid, nu); m.actuator_dyntype = upi(s, h->actuator_dyntype, nu);
  m.actuator_biastype = upi(s, h->actuator_biastype, nu); m.actuator_ctrllimited = upi(s, h->actuator_ctrllimited, nu);
  m.actuator_forcelimited = upi(s, h->actuator_forcelimited, nu); m.actuator_actadr = upi(s, h->actuator_actadr, nu);
  m.actuator_dynprm = upf(s, h->actuator_dynprm, 3 * nu); m.actuator_gainprm = upf(s, h->actuator_gainprm, 3 * nu); m.actuator_biasprm = upf(s, h->actuator_biasprm, 3 * nu);
  m.actuator_ctrlrange = upf(s, h->actuator_ctrlrange, 2 * nu); m.actuator_forcerange = upf(s, h->actuator_forcerange, 2 * nu);
  m.sensor_type = upi(s, h->sensor_type, m.nsensor); m.sensor_objid = upi(s, h->sensor_objid, m.nsensor);
  m.sensor_adr = upi(s, h->sensor_adr, m.nsensor); m.sensor_dim = upi(s, h->sensor_dim, m.nsensor);
  s->h_dof_parent.assign(h->dof_parentid, h->dof_parentid + nv); s->h_dof_Madr.assign(h->dof_Madr, h->dof_Madr + nv);
  s->h_qpos0.assign(h->qpos0, h->qpos0 + m.nq);
  s->h_body_lastdof.assign(h->body_lastdof, h->body_lastdof + nb); s->h_geom_bodyid.assign(h->geom_bodyid, h->geom_bodyid + ng);
  return 0;
}

static int alloc_data(FbSim* s, int N) {
  DevData& d = s->d; const DevModel& m = s->m;
  memset(&d, 0, sizeof(d));
  d.N = N; d.Np = (N + FB_WPB - 1) / FB_WPB * FB_WPB; d.sens_mode = -1;
  // pass 1: lay the per-env record out (offsets in 4-byte slots); pass 2: one allocation, rebase the pointers
  std::vector<std::pair<void**, size_t>> fields;
  size_t off = 0;
#define FA(field, n) { off = (off + 3) & ~(size_t)3; fields.push_back({(void**)&d.field, off}); off += (size_t)(n); }   // arrays start on 16-byte boundaries
#define IA(field, n) FA(field, n)
  FA(qpos, m.nq) FA(qvel, m.nv) FA(act, m.na + 1) FA(ctrl, m.nu + 1) FA(qacc, m.nv) FA(dof_isd, m.nv) FA(time, 1)
  // vectors / matrices / inertias padded to whole float4s (fb_math.h: FB_V3S ...)
  FA(ref, 3) FA(xpos, FB_V3S * m.nbody) FA(xquat, 4 * m.nbody) FA(xmat, FB_M3S * m.nbody) FA(xipos, FB_V3S * m.nbody) FA(ximat, FB_M3S * m.nbody)
  FA(geom_xpos, FB_V3S * m.ngeom) FA(geom_xmat, FB_M3S * m.ngeom) FA(site_xpos, FB_V3S * (m.nsite + 1)) FA(site_xmat, FB_M3S * (m.nsite + 1))
  FA(Sang, FB_V3S * m.nv) FA(Slin, FB_V3S * m.nv) FA(inert10, FB_I10S * m.nbody) FA(crb10, FB_I10S * m.nbody)
  FA(qM, m.nM) FA(qLD, m.nM) FA(qLDe, m.nM)
  FA(bvel, FB_S6S * m.nbody) FA(bacc, FB_S6S * m.nbody) FA(bfrc, FB_S6S * m.nbody) FA(bfl, FB_S6S * m.nbody) FA(bfrc0, FB_S6S * m.nbody) FA(bdel, FB_S6S * m.nbody)
  FA(qfrc_bias, m.nv) FA(qfrc_passive, m.nv) FA(qfrc_actuator, m.nv) FA(qfrc_smooth, m.nv) FA(qfrc_zf, m.nv) FA(qfrc_constraint, m.nv) FA(qtmp, m.nv)
  FA(act_dot, m.na + 1) FA(actuator_force, m.nu + 1)
  IA(ncon, 1) FA(con_dist, FB_MAXCON) FA(con_pos, 3 * FB_MAXCON) FA(con_frame, 9 * FB_MAXCON) IA(con_geom1, FB_MAXCON) IA(con_geom2, FB_MAXCON)
  IA(con_efcadr, FB_MAXCON) IA(con_dim, FB_MAXCON) FA(con_mu, FB_MAXCON) FA(con_fric, 2 * FB_MAXCON)
  FA(tmp_con, 13 * 4 * FB_MAXCAND) IA(tmp_geom, 2 * 4 * FB_MAXCAND)
  IA(nefc, 1) IA(efc_type, FB_MAXEFC) IA(efc_id, FB_MAXEFC)
  FA(efc_pos, FB_MAXEFC) FA(efc_margin, FB_MAXEFC) FA(efc_D, FB_MAXEFC) FA(efc_R, FB_MAXEFC) FA(efc_K, FB_MAXEFC) FA(efc_B, FB_MAXEFC)
  FA(efc_imp, FB_MAXEFC) FA(efc_aref, FB_MAXEFC) FA(efc_b, FB_MAXEFC) FA(efc_force, FB_MAXEFC) FA(efc_jarws, FB_MAXEFC)
  IA(efc_la, FB_MAXEFC) IA(efc_lb, FB_MAXEFC) IA(efc_key, FB_MAXEFC) IA(prev_key, FB_MAXEFC) IA(prev_n, 1) FA(prev_lam, FB_MAXEFC)
  FA(sensordata, m.nsensordata + 1) FA(sensor_sum, m.nsensordata + 1) IA(flags, 1) IA(niter, 1) IA(hold, 1)
  // large, sparsely touched arrays last
  FA(efc_w, (size_t)S_NSLOT * FB_MAXEFC)
  FA(efc_A, (size_t)FB_MAXEFC * (FB_MAXEFC + 1) / 2) FA(efc_G, (size_t)FB_MAXEFC * (FB_MAXEFC + 1) / 2)
  FA(efc_J, (size_t)FB_MAXEFC * FB_JROW) FA(efc_Z, (size_t)FB_MAXEFC * FB_JROW)
#undef FA
#undef IA
  off = (off + 31) & ~(size_t)31;              // records start on 128-byte boundaries
  if ((double)off * d.Np >= 4.0e9) { s->err = "too many envs for 32-bit record indexing"; return -5; }
  d.rec = (unsigned)off;
  float* base = dalloc<float>(s, (size_t)d.Np * off);
  if (!base) { s->err = "out of device memory (env records)"; return -4; }
  for (auto& f : fields) *f.first = (void*)(base + f.second);
#ifdef FB_CLK
  d.clk = (long long*)dalloc<long long>(s, 32 * 4096);
#endif
#ifndef FB_EMU
  // FB_HEAVY_KERNEL=1: queue of envs for the heavy-env solve kernel (fb_run_solve_big).  Measured on the B200 and left OFF: with 2 of 4096
  // envs above 32 rows the solve stage went from 2.77 to 4.04 ms per control step -- a single warp needs > 100 us for a 48-row problem even
  // from shared memory, and behind the main kernel that time is serial, whereas inline (generic code on the env's global record) it
  // overlaps with the other envs' solves and only stretches the kernel's tail by ~15 %.
  if (s->fuse == 0 && s->split == 1 && getenv("FB_HEAVY_KERNEL") && atoi(getenv("FB_HEAVY_KERNEL"))) { d.heavy_count = dalloc<int>(s, 4); d.heavy_list = dalloc<int>(s, d.Np); }
#endif
  d.obs_dim = m.nq + m.nv + m.na + 2 * m.nsensordata + 12 + 3 * m.nsite + 3;
  d.obs = dalloc<float>(s, (size_t)d.obs_dim * d.Np);
  s->stage_cap = 0; s->stage = nullptr; s->stage_i = nullptr;
  return 0;
}

// host <-> device transfers of one field: host side is AoS [N][n]; device side is n slots of every record
static void field_to_host(FbSim* s, const void* dev, int n, void* dst) {
  if (n <= 0) return;
#ifndef FB_EMU
  cudaMemcpy2DAsync(dst, (size_t)n * 4, dev, (size_t)s->d.rec * 4, (size_t)n * 4, s->d.N, cudaMemcpyDeviceToHost, s->stream);
  cudaStreamSynchronize(s->stream);
#else
  for (int e = 0; e < s->d.N; e++) memcpy((char*)dst + (size_t)e * n * 4, (const char*)dev + (size_t)e * s->d.rec * 4, (size_t)n * 4);
#endif
}
// padded device array (count elements of `width` floats every `stride` slots) -> compact host array [N][count * width]
static void field_to_host_strided(FbSim* s, const void* dev, int count, int width, int stride, float* dst) {
  std::vector<float> tmp((size_t)s->d.N * count * stride);
  field_to_host(s, dev, count * stride, tmp.data());
  for (int e = 0; e < s->d.N; e++) for (int i = 0; i < count; i++) for (int k = 0; k < width; k++)
    dst[((size_t)e * count + i) * width + k] = tmp[((size_t)e * count + i) * stride + k];
}
static void field_from_host(FbSim* s, void* dev, int n, const void* src) {
  if (n <= 0) return;
#ifndef FB_EMU
  // stream-ordered (the handle's stream is non-blocking: legacy default-stream copies would not order with its kernels)
  cudaMemcpy2DAsync(dev, (size_t)s->d.rec * 4, src, (size_t)n * 4, (size_t)n * 4, s->d.N, cudaMemcpyHostToDevice, s->stream);
  for (int e = s->d.N; e < s->d.Np; e++) cudaMemcpyAsync((char*)dev + (size_t)e * s->d.rec * 4, src, (size_t)n * 4, cudaMemcpyHostToDevice, s->stream);   // pad envs mirror env 0
  cudaStreamSynchronize(s->stream);
#else
  for (int e = 0; e < s->d.Np; e++) memcpy((char*)dev + (size_t)e * s->d.rec * 4, (const char*)src + (size_t)(e < s->d.N ? e : 0) * n * 4, (size_t)n * 4);
#endif
}
// device staging buffer for scatter-style uploads (ctrl, ghost pose, partial resets)
static int ensure_stage(FbSim* s, size_t floats, size_t ints) {
#ifndef FB_EMU
  if (floats > s->stage_cap || ints > s->stage_icap) cudaDeviceSynchronize();
#endif
  if (floats > s->stage_cap) { s->stage = dalloc<float>(s, floats); s->stage_cap = floats; if (!s->stage) return -4; }
  if (ints > s->stage_icap) { s->stage_i = dalloc<int>(s, ints); s->stage_icap = ints; if (!s->stage_i) return -4; }
  return 0;
}
static void upload_async(FbSim* s, void* dst, const void* src, size_t bytes) {
#ifndef FB_EMU
  cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s->stream);
#else
  memcpy(dst, src, bytes);
#endif
}

static int sync_stream(FbSim* s) {
#ifndef FB_EMU
  FB_CUDA_OK(cudaStreamSynchronize(s->stream));
  FB_CUDA_OK(cudaGetLastError());
#endif
  return 0;
}

extern "C" {

const char* fb_version(void) {
#ifdef FB_EMU
  return "flybody_b200 0.2 (host emulation build: tests only)";
#else
  return "flybody_b200 0.2 (sm_100a)";
#endif
}

int fb_create(const FbModel* hm, int n_envs, int device, FbHandle* out) {
  if (!hm || !out || n_envs <= 0) return -1;
  FbSim* s = new FbSim();
  s->device = device; s->launches = 0; s->last_ms = 0; s->hm = *hm; s->first_substep = 1; s->hold_pending = 0; s->prof_on = 0; memset(s->prof_ms, 0, sizeof(s->prof_ms)); memset(s->prof_n, 0, sizeof(s->prof_n));
  s->fuse = getenv("FB_FUSE") ? atoi(getenv("FB_FUSE")) : FB_FUSE_DEFAULT; if (s->fuse < 0 || s->fuse > 6 || s->fuse == 5) s->fuse = FB_FUSE_DEFAULT;
  s->ref_slots = nullptr; s->ref_slot_len = 0; s->eye_out = nullptr; s->hfield_dev = nullptr; s->hmax_dev = nullptr; s->cmax_dev = nullptr; s->hf_nbr = s->hf_nbc = 0; s->bank_dev = s->bank_hmax_dev = s->bank_cmax_dev = nullptr; s->bank_n = 0; s->eye_bytes = 0; s->hf_on = false; s->hf_nrow = s->hf_ncol = 0;
  s->op_step_dev = nullptr; s->op_first_dev = nullptr; s->stage_cap = 0; s->stage_icap = 0; s->stage = nullptr; s->stage_i = nullptr;
#ifndef FB_EMU
  if (cudaSetDevice(device) != cudaSuccess) { delete s; return -2; }
  cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  { int n_sm = 0; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device); s->n_sm = n_sm > 0 ? n_sm : 148; }
  cudaEventCreate(&s->ev0); cudaEventCreate(&s->ev1);
  s->graphs_on = getenv("FB_NO_GRAPH") == nullptr;
  s->split = getenv("FB_SPLIT") ? atoi(getenv("FB_SPLIT")) : FB_SPLIT_DEFAULT; if (s->split < 1 || s->split > 4) s->split = 1;
  s->chain_stagger = getenv("FB_STAGGER") ? atoi(getenv("FB_STAGGER")) : (s->split == 2 ? 3 : 2);
  s->chain_count = -1; s->chain_idx = 0; s->cur_stream = s->stream;
  for (int i = 0; i < 3; i++) cudaStreamCreateWithFlags(&s->aux[i], cudaStreamNonBlocking);
  for (int i = 0; i < 4; i++) { cudaEventCreateWithFlags(&s->chain_ev[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&s->join_ev[i], cudaEventDisableTiming); }
  cudaEventCreateWithFlags(&s->fork_ev, cudaEventDisableTiming); cudaEventCreateWithFlags(&s->vel_ev, cudaEventDisableTiming);
  s->overlap_vel = getenv("FB_OVERLAP_VEL") ? atoi(getenv("FB_OVERLAP_VEL")) : FB_OVERLAP_VEL_DEFAULT;
  // small batches are latency-bound (a per-CTA copy of the sweep program in shared memory halves the sweeps' latency);
  // at full occupancy the copy costs more than it saves and the compact program is read in place.  FB_BLOB=0/1 overrides.
  s->blob_in_smem = getenv("FB_BLOB") ? atoi(getenv("FB_BLOB")) : (n_envs <= 1024);
#endif
  int rc = build_model(s, hm);
  if (rc == 0) rc = alloc_data(s, n_envs);
  s->cur_e0 = 0; s->cur_n = s->d.Np;
  *out = s;
  if (rc != 0) return rc;
#ifndef FB_EMU
  cudaDeviceSynchronize();                 // model uploads / memsets ran on the legacy stream
#endif
  std::vector<float> q0((size_t)n_envs * hm->nq);
  for (int e = 0; e < n_envs; e++) for (int i = 0; i < hm->nq; i++) q0[(size_t)e * hm->nq + i] = (float)hm->qpos0[i];
  field_from_host(s, s->d.qpos, hm->nq, q0.data());
  return fb_forward(s);
}

int fb_destroy(FbHandle s) {
  if (!s) return -1;
#ifndef FB_EMU
  cudaSetDevice(s->device); cudaStreamSynchronize(s->stream);
  for (auto& g : s->graph) if (g.exec) cudaGraphExecDestroy(g.exec);
  cudaEventDestroy(s->ev0); cudaEventDestroy(s->ev1); cudaStreamDestroy(s->stream);
#endif
  for (void* p : s->allocs) dev_free(p);
  delete s;
  return 0;
}

#ifdef FB_CLK
extern "C" int fb_clk_read(FbHandle s, long long* dst) { if (!s) return -1; cudaStreamSynchronize(s->stream); cudaMemcpy(dst, s->d.clk, sizeof(long long) * 32 * 4096, cudaMemcpyDeviceToHost); return (int)(s->launches % 4096); }
#endif
#ifdef FB_EMU
extern "C" int fb_emu_convex_hfield(int type, const float* gp, const float* gm, const float* gs, float margin, const float* hp, const float* hm,
                                    const float* hf_size, int nrow, int ncol, const float* data, float* out /* [max][7] */, int max) {
  HfCon c[64]; M3 Gm, Hm; for (int k = 0; k < 9; k++) { Gm.m[k] = gm[k]; Hm.m[k] = hm[k]; }
  if (max > 64) max = 64;
  int n = col_convex_hfield(c, max, margin, type, v3(gp[0], gp[1], gp[2]), Gm, v3(gs[0], gs[1], gs[2]), v3(hp[0], hp[1], hp[2]), Hm, hf_size, nrow, ncol, data);
  for (int k = 0; k < n; k++) { out[7 * k] = c[k].dist; out[7 * k + 1] = c[k].pos.x; out[7 * k + 2] = c[k].pos.y; out[7 * k + 3] = c[k].pos.z; out[7 * k + 4] = c[k].n.x; out[7 * k + 5] = c[k].n.y; out[7 * k + 6] = c[k].n.z; }
  return n;
}
extern "C" void fb_emu_convex_stats(long* out) { for (int i = 0; i < 4; i++) { out[i] = g_convex_stats[i]; g_convex_stats[i] = 0; } }
// host-emulation build only (tests): the fp32 generic-convex narrowphase on one pair; out = dist, pos[3], normal[3]
extern "C" int fb_emu_convex_pair(int t1, const float* p1, const float* m1, const float* s1, int t2, const float* p2, const float* m2, const float* s2,
                                  float margin, float* out) {
  M3 R1, R2; for (int k = 0; k < 9; k++) { R1.m[k] = m1[k]; R2.m[k] = m2[k]; }
  RawCon c; int n = col_convex(&c, margin, t1, v3(p1[0], p1[1], p1[2]), R1, v3(s1[0], s1[1], s1[2]), t2, v3(p2[0], p2[1], p2[2]), R2, v3(s2[0], s2[1], s2[2]));
  if (n) { out[0] = c.dist; out[1] = c.pos.x; out[2] = c.pos.y; out[3] = c.pos.z; out[4] = c.n.x; out[5] = c.n.y; out[6] = c.n.z; }
  return n;
}
#endif
const char* fb_last_error(FbHandle s) { return s ? s->err.c_str() : "null handle"; }
int fb_n_envs(FbHandle s) { return s ? s->d.N : -1; }
int fb_n_envs_padded(FbHandle s) { return s ? s->d.Np : -1; }
int fb_record_stride(FbHandle s) { return s ? (int)s->d.rec : -1; }
long long fb_launch_count(FbHandle s) { return s ? s->launches : -1; }
float fb_last_step_ms(FbHandle s) {
#ifndef FB_EMU
  if (!s) return -1; float ms = 0; cudaEventSynchronize(s->ev1); cudaEventElapsedTime(&ms, s->ev0, s->ev1); return ms;
#else
  return 0;
#endif
}
void* fb_stream(FbHandle s) {
#ifndef FB_EMU
  return s ? (void*)s->stream : nullptr;
#else
  return nullptr;
#endif
}
int fb_sync(FbHandle s) { if (!s) return -1; return sync_stream(s); }
int fb_set_solver(FbHandle s, float tolerance, int max_iter) {
  if (!s) return -1;
  if (tolerance > 0) s->m.tolerance = tolerance;
  if (max_iter > 0) s->m.max_iter = max_iter;
  return 0;
}

int fb_forward(FbHandle s) {
  if (!s) return -1;
#ifndef FB_EMU
  cudaSetDevice(s->device);
#endif
  launch_step1(s);
  launch_step2(s, false);
  return sync_stream(s);
}

// the launch sequence of one control step: n x [ step2 ; step1 ], sensor accumulation restarted at the first substep
static void step_sequence(FbSim* s, int n_substeps) {
  if (s->fuse == 3) {                       // the whole control step in one launch
    s->d.do_integrate = 1; s->d.sens_mode = -1;
    fb_launch_fused<GSmooth, GrpSolve, GFinish, GPos, GCol, GProj, GVel>(s, K_STEP, n_substeps, 1);
  }
#ifndef FB_EMU
  else if (s->fuse == 0 && s->split > 1) {
    // env ranges as staggered chains on their own streams (captured into the step graph as parallel branches)
    const int S = s->split, per = ((s->d.Np / FB_WPB + S - 1) / S) * FB_WPB;
    for (int h = 0; h < S; h++) {
      s->cur_stream = h == 0 ? s->stream : s->aux[h - 1];
      s->cur_e0 = h * per; s->cur_n = std::min(per, s->d.Np - h * per);
      if (s->cur_n <= 0) { s->cur_n = 0; continue; }
      if (h > 0) cudaStreamWaitEvent(s->cur_stream, s->chain_ev[h - 1], 0);
      s->chain_idx = h; s->chain_count = 0;
      for (int k = 0; k < n_substeps; k++) {
        launch_step2(s, true);
        s->d.sens_mode = (k == 0) ? 1 : 0;
        launch_step1(s);
        s->d.sens_mode = -1;
      }
      s->chain_count = -1;
      if (h > 0) { cudaEventRecord(s->join_ev[h], s->cur_stream); cudaStreamWaitEvent(s->stream, s->join_ev[h], 0); }
    }
    s->cur_stream = s->stream; s->cur_e0 = 0; s->cur_n = s->d.Np;
  }
#endif
  else for (int k = 0; k < n_substeps; k++) {
    if (s->fuse == 0) {
      launch_step2(s, true);
      s->d.sens_mode = (k == 0) ? 1 : 0;
      launch_step1(s);
    } else if (s->fuse == 4 || s->fuse == 6) {
      // only the straggler-prone stage pairs share a launch: the Newton solve (1-7 iterations per env) with the finish stage,
      // and (6) collision (MPR tail on one warp per block) with the constraint rows
      s->d.do_integrate = 1;
      fb_launch<ShTree, FB_ST_SMOOTH>(s, K_SMOOTH, dyn_tsolve(s->m), -1, s->blob_in_smem ? s->m.ts_blob_words : 0);
      fb_launch_fused<GrpSolve, GFinish>(s, K_STEP2, 1, 0);
      s->d.sens_mode = (k == 0) ? 1 : 0;
      fb_launch<ShTree, FB_ST_POS>(s, K_POS, dyn_pos(s->m));
      if (s->fuse == 6) fb_launch_fused<GCol, GProj>(s, K_STEP1, 1, 0);
      else { fb_launch<ShCol, FB_ST_COL>(s, K_COL, dyn_col(s->m)); fb_launch<ShCon, FB_ST_PROJ>(s, K_PROJ, dyn_proj(s->m)); }
      fb_launch<ShTree, FB_ST_VEL>(s, K_VEL, dyn_vel(s->m));
    } else if (s->fuse == 1) {
      s->d.do_integrate = 1;
      fb_launch_fused<GSmooth, GrpSolve, GFinish>(s, K_STEP2, 1, 0);
      s->d.sens_mode = (k == 0) ? 1 : 0;
      fb_launch_fused<GPos, GCol, GProj, GVel>(s, K_STEP1, 1, 0);
    } else {
      s->d.do_integrate = 1; s->d.sens_mode = (k == 0) ? 1 : 0;       // only the last group (vel) reads sens_mode
      fb_launch_fused<GSmooth, GrpSolve, GFinish, GPos, GCol, GProj, GVel>(s, K_STEP, 1, 0);
    }
    s->d.sens_mode = -1;
  }
  s->d.nsub_done = n_substeps;
  if (s->hold_pending) { fb_launch<ShNone, Ph<ph_clear_hold>>(s, K_MISC); s->hold_pending = 0; }
}
#ifndef FB_EMU
// Kernel parameters (DevModel, DevData by value) are frozen into a captured graph, so a graph is reused only while the
// parts of the handle state that the step kernels read are unchanged; staging pointers of scatter / reset launches are not
// read by them and are masked out of the comparison.
static DevData canonical(const DevData& d) {
  DevData c = d;
  c.sc_field = nullptr; c.sc_idx = nullptr; c.sc_vals = nullptr; c.sc_k = 0; c.sc_nan0 = 0;
  c.rst_ids = nullptr; c.rst_qpos = nullptr; c.rst_qvel = nullptr; c.rst_n = 0; c.rst_has_qvel = 0; c.rst_hold = 0;
  return c;
}
#endif
int fb_step(FbHandle s, int n_substeps) {
  if (!s || n_substeps <= 0) return -1;
#ifndef FB_EMU
  cudaSetDevice(s->device);
  cudaEventRecord(s->ev0, s->stream);
  // The 7 n launches of a control step are replayed from a CUDA graph: captured the second time the same (n_substeps,
  // hold) signature is stepped with unchanged handle state, so the first call still runs (and configures) plain launches.
  const int hold = s->hold_pending ? 1 : 0;
  bool done = false;
  if (s->graphs_on && !s->prof_on) {
    DevData c = canonical(s->d);
    FbSim::StepGraph& g = s->graph[hold];
    const bool same = g.seen && g.n_sub == n_substeps && memcmp(&c, &g.d, sizeof(DevData)) == 0 && memcmp(&s->m, &g.m, sizeof(DevModel)) == 0;
    if (same && g.exec) {
      if (cudaGraphLaunch(g.exec, s->stream) == cudaSuccess) {
        s->d.do_integrate = 1; s->d.sens_mode = -1; s->d.nsub_done = n_substeps; s->hold_pending = 0; s->launches += g.launches;
        done = true;
      } else { cudaGetLastError(); cudaGraphExecDestroy(g.exec); g.exec = nullptr; g.seen = false; }
    } else if (same && !g.failed) {
      long long l0 = s->launches; cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
        step_sequence(s, n_substeps);
        cudaError_t e1 = cudaStreamEndCapture(s->stream, &graph);
        if (e1 == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess && cudaGraphLaunch(g.exec, s->stream) == cudaSuccess) {
          g.launches = s->launches - l0; done = true;
        } else { cudaGetLastError(); g.exec = nullptr; g.failed = true; s->launches = l0; s->hold_pending = hold; }
        if (graph) cudaGraphDestroy(graph);
      } else { cudaGetLastError(); g.failed = true; }
    } else {
      if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
      g.seen = true; g.n_sub = n_substeps; g.d = c; g.m = s->m;
    }
  }
  if (!done) step_sequence(s, n_substeps);
  cudaEventRecord(s->ev1, s->stream);
  if (cudaGetLastError() != cudaSuccess) { s->err = "kernel launch failed"; return -2; }
#else
  step_sequence(s, n_substeps);
#endif
  return 0;
}

static void* field_ptr(FbSim* s, int field, int* n) {
  const DevModel& m = s->m; DevData& d = s->d;
  switch (field) {
    case FB_QPOS: *n = m.nq; return d.qpos;
    case FB_QVEL: *n = m.nv; return d.qvel;
    case FB_ACT: *n = m.na; return d.act;
    case FB_CTRL: *n = m.nu; return d.ctrl;
    case FB_QACC: *n = m.nv; return d.qacc;
    case FB_QACC_WARMSTART: *n = m.nv; return d.qacc;        /* the dual solver warm-starts from forces; kept for ABI compatibility */
    case FB_SENSORDATA: *n = m.nsensordata; return d.sensordata;
    case FB_SENSOR_MEAN: *n = m.nsensordata; return d.sensor_sum;
    case FB_XPOS: *n = 3 * m.nbody; return d.xpos;            /* padded on the device (fb_math.h): host reads go through field_to_host_strided */
    case FB_XMAT: *n = 9 * m.nbody; return d.xmat;
    case FB_SITE_XPOS: *n = 3 * m.nsite; return d.site_xpos;
    case FB_SITE_XMAT: *n = 9 * m.nsite; return d.site_xmat;
    case FB_TIME: *n = 1; return d.time;
    case FB_QFRC_SMOOTH: *n = m.nv; return d.qfrc_smooth;
    case FB_QFRC_CONSTRAINT: *n = m.nv; return d.qfrc_constraint;
    case FB_QFRC_PASSIVE: *n = m.nv; return d.qfrc_passive;
    case FB_QFRC_BIAS: *n = m.nv; return d.qfrc_bias;
    case FB_QFRC_ACTUATOR: *n = m.nv; return d.qfrc_actuator;
    case FB_EFC_FORCE: *n = FB_MAXEFC; return d.efc_force;
    default: *n = -1; return nullptr;
  }
}

int fb_field_size(FbHandle s, int field) {
  if (!s) return -1;
  int n; if (field_ptr(s, field, &n)) return n;
  switch (field) {
    case FB_SUBTREE_COM: return 3 * s->m.nbody;
    case FB_NCON: case FB_NEFC: case FB_SOLVER_NITER: case FB_FLAGS: return 1;
    case FB_QM_DENSE: return s->m.nv * s->m.nv;
    case FB_CONTACT: return 16 * FB_MAXCON;
    default: return -1;
  }
}

int fb_get(FbHandle s, int field, void* dst, int is_device) {
  if (!s || !dst) return -1;
  if (sync_stream(s) != 0) return -2;
  int n; void* p = field_ptr(s, field, &n);
  const DevModel& m = s->m; int N = s->d.N;
  if (is_device) { if (!p) { s->err = "field has no device array"; return -1; } *(void**)dst = p; return 0; }
  float* out = (float*)dst;
  if (p) {
    if (field == FB_XPOS) field_to_host_strided(s, p, m.nbody, 3, FB_V3S, out);
    else if (field == FB_XMAT) field_to_host_strided(s, p, m.nbody, 9, FB_M3S, out);
    else if (field == FB_SITE_XPOS) field_to_host_strided(s, p, m.nsite, 3, FB_V3S, out);
    else if (field == FB_SITE_XMAT) field_to_host_strided(s, p, m.nsite, 9, FB_M3S, out);
    else field_to_host(s, p, n, out);
    if (field == FB_XPOS || field == FB_SITE_XPOS) {     // stored relative to ref
      std::vector<float> ref((size_t)3 * N); field_to_host(s, s->d.ref, 3, ref.data());
      for (int e = 0; e < N; e++) for (int i = 0; i < n; i++) out[(size_t)e * n + i] += ref[(size_t)e * 3 + i % 3];
    }
    if (field == FB_SENSOR_MEAN && s->d.nsub_done > 0) for (size_t i = 0; i < (size_t)N * n; i++) out[i] /= (float)s->d.nsub_done;
    return 0;
  }
  std::vector<int> iv(N);
  switch (field) {
    case FB_NCON: field_to_host(s, s->d.ncon, 1, iv.data()); for (int e = 0; e < N; e++) out[e] = (float)iv[e]; return 0;
    case FB_NEFC: field_to_host(s, s->d.nefc, 1, iv.data()); for (int e = 0; e < N; e++) out[e] = (float)iv[e]; return 0;
    case FB_SOLVER_NITER: field_to_host(s, s->d.niter, 1, iv.data()); for (int e = 0; e < N; e++) out[e] = (float)iv[e]; return 0;
    case FB_FLAGS: field_to_host(s, s->d.flags, 1, iv.data()); for (int e = 0; e < N; e++) out[e] = (float)iv[e]; return 0;
    case FB_SUBTREE_COM: {
      std::vector<float> crb((size_t)10 * m.nbody * N), ref((size_t)3 * N);
      field_to_host_strided(s, s->d.crb10, m.nbody, 10, FB_I10S, crb.data()); field_to_host(s, s->d.ref, 3, ref.data());
      for (int e = 0; e < N; e++) for (int b = 0; b < m.nbody; b++) { const float* c = &crb[((size_t)e * m.nbody + b) * 10];
        for (int i = 0; i < 3; i++) out[((size_t)e * m.nbody + b) * 3 + i] = (c[0] > 0 ? c[1 + i] / c[0] : 0.0f) + ref[(size_t)e * 3 + i]; }
      return 0; }
    case FB_QM_DENSE: {
      std::vector<float> qm((size_t)m.nM * N); field_to_host(s, s->d.qM, m.nM, qm.data());
      int nv = m.nv; memset(out, 0, sizeof(float) * (size_t)N * nv * nv);
      for (int e = 0; e < N; e++) for (int i = 0; i < nv; i++) { int t = 0; for (int j = i; j >= 0; j = s->h_dof_parent[j], t++) {
        float v = qm[(size_t)e * m.nM + s->h_dof_Madr[i] + t]; out[((size_t)e * nv + i) * nv + j] = v; out[((size_t)e * nv + j) * nv + i] = v; } }
      return 0; }
    case FB_CONTACT: {
      std::vector<float> dist((size_t)FB_MAXCON * N), pos((size_t)3 * FB_MAXCON * N), frame((size_t)9 * FB_MAXCON * N), mu((size_t)FB_MAXCON * N), ref((size_t)3 * N);
      std::vector<int> g1((size_t)FB_MAXCON * N), g2((size_t)FB_MAXCON * N), ea((size_t)FB_MAXCON * N), dm((size_t)FB_MAXCON * N), nc(N);
      field_to_host(s, s->d.con_dist, FB_MAXCON, dist.data()); field_to_host(s, s->d.con_pos, 3 * FB_MAXCON, pos.data());
      field_to_host(s, s->d.con_frame, 9 * FB_MAXCON, frame.data()); field_to_host(s, s->d.con_mu, FB_MAXCON, mu.data()); field_to_host(s, s->d.ref, 3, ref.data());
      field_to_host(s, s->d.con_geom1, FB_MAXCON, g1.data()); field_to_host(s, s->d.con_geom2, FB_MAXCON, g2.data());
      field_to_host(s, s->d.con_efcadr, FB_MAXCON, ea.data()); field_to_host(s, s->d.con_dim, FB_MAXCON, dm.data()); field_to_host(s, s->d.ncon, 1, nc.data());
      memset(out, 0, sizeof(float) * (size_t)N * 16 * FB_MAXCON);
      for (int e = 0; e < N; e++) for (int c = 0; c < nc[e] && c < FB_MAXCON; c++) {
        float* o = out + ((size_t)e * FB_MAXCON + c) * 16; size_t ec = (size_t)e * FB_MAXCON + c;
        o[0] = dist[ec];
        for (int i = 0; i < 3; i++) { o[1 + i] = pos[ec * 3 + i] + ref[(size_t)e * 3 + i]; o[4 + i] = frame[ec * 9 + i]; }
        o[7] = (float)g1[ec]; o[8] = (float)g2[ec]; o[9] = (float)dm[ec];
        o[12] = (float)ea[ec]; o[10] = o[12] >= 0 ? 1.0f : 0.0f; o[11] = mu[ec];
      }
      return 0; }
    default: s->err = "unknown field"; return -1;
  }
}

int fb_set(FbHandle s, int field, const float* src) {
  if (!s || !src) return -1;
  if (sync_stream(s) != 0) return -2;
  int n; void* p = field_ptr(s, field, &n);
  if (!p || !(field == FB_QPOS || field == FB_QVEL || field == FB_ACT || field == FB_CTRL || field == FB_QACC_WARMSTART || field == FB_QACC || field == FB_TIME)) { s->err = "field is not writable"; return -1; }
  if (n > 0) field_from_host(s, p, n, src);
  return 0;
}

// scatter vals[N][k] (already on the device, row-major) into columns idx[k] (device ints, or NULL = 0..k-1) of `field`
static void launch_scatter(FbSim* s, float* field, const int* idx_dev, const float* vals_dev, int k, int nan_to_zero = 0) {
  s->d.sc_field = field; s->d.sc_idx = idx_dev; s->d.sc_vals = vals_dev; s->d.sc_k = k; s->d.sc_nan0 = nan_to_zero;
  fb_launch<ShNone, Ph<ph_scatter>>(s, K_MISC);
}

int fb_set_ctrl(FbHandle s, const float* ctrl, int is_device) {
  if (!s || !ctrl) return -1;
  const float* src = ctrl;
  const int k = s->act_map_dev ? s->n_action : s->m.nu;     // row length: actions (mapped on the device) or ctrl
  if (!is_device) {
    if (ensure_stage(s, (size_t)s->d.N * std::max(k, s->m.nu), 0) != 0) { s->err = "out of device memory (staging)"; return -4; }
    upload_async(s, s->stage, ctrl, sizeof(float) * (size_t)s->d.N * k);
    src = s->stage;
  }
  launch_scatter(s, s->d.ctrl, s->act_map_dev, src, k, s->act_map_dev ? 1 : 0);
  return 0;
}
int fb_set_action_map(FbHandle s, const int32_t* ctrl_index, int n_action) {
  if (!s) return -1;
  if (!ctrl_index || n_action <= 0) { s->act_map_dev = nullptr; s->n_action = 0; return 0; }      // back to plain ctrl rows
  for (int c = 0; c < n_action; c++) if (ctrl_index[c] >= s->m.nu || ctrl_index[c] < -1) { s->err = "fb_set_action_map: ctrl index out of range (valid: -1 = no target, 0..nu-1)"; return -1; }
  if (sync_stream(s) != 0) return -2;
  std::vector<int> v(ctrl_index, ctrl_index + n_action);
  s->act_map_dev = up(s, v); s->n_action = n_action;
  return 0;
}

int fb_write_state(FbHandle s, int field, const int32_t* idx, int k, const float* vals) {
  if (!s || !idx || !vals || k <= 0) return -1;
  int n; float* p = (float*)field_ptr(s, field, &n);
  if (!p || !(field == FB_QPOS || field == FB_QVEL || field == FB_ACT)) { s->err = "fb_write_state: field must be qpos, qvel or act"; return -1; }
  for (int c = 0; c < k; c++) if (idx[c] < 0 || idx[c] >= n) { s->err = "fb_write_state: index out of range"; return -1; }
  // separate staging regions so that several writes can be in flight before the next sync
  size_t need = (size_t)s->d.N * s->m.nu + (size_t)4 * s->d.N * 16;
  if (k > 16) return -1;
  if (ensure_stage(s, need, 64) != 0) { s->err = "out of device memory (staging)"; return -4; }
  int slot = s->ws_slot++ & 3;
  float* vdev = s->stage + (size_t)s->d.N * s->m.nu + (size_t)slot * s->d.N * 16;
  int* idev = s->stage_i + slot * 16;
  upload_async(s, idev, idx, sizeof(int) * k);
  upload_async(s, vdev, vals, sizeof(float) * (size_t)s->d.N * k);
  launch_scatter(s, p, idev, vdev, k);
  return 0;
}

static int do_reset(FbSim* s, const int32_t* env_ids, int n, const float* qpos, const float* qvel, int hold) {
  const DevModel& m = s->m; int N = s->d.N, Np = s->d.Np;
  if (sync_stream(s) != 0) return -2;
  std::vector<int> ids; std::vector<float> qp, qv;
  bool all = (env_ids == nullptr);
  if (all) n = N;
  for (int k = 0; k < n; k++) { int e = all ? k : env_ids[k]; if (e < 0 || e >= N) { s->err = "reset: env id out of range"; return -1; } ids.push_back(e); }
  qp.resize((size_t)ids.size() * m.nq);
  for (size_t k = 0; k < ids.size(); k++) for (int i = 0; i < m.nq; i++) qp[k * m.nq + i] = qpos ? qpos[k * m.nq + i] : (float)s->h_qpos0[i];
  if (qvel) qv.assign(qvel, qvel + (size_t)ids.size() * m.nv);
  if (all) for (int e = N; e < Np; e++) {          // pad envs mirror env 0
    ids.push_back(e); qp.insert(qp.end(), qp.begin(), qp.begin() + m.nq); if (qvel) qv.insert(qv.end(), qv.begin(), qv.begin() + m.nv);
  }
  int cnt = (int)ids.size();
  if (cnt > s->rst_cap) {
    s->rst_cap = std::max(cnt, Np);
    s->rst_ids_dev = dalloc<int>(s, s->rst_cap); s->rst_qpos_dev = dalloc<float>(s, (size_t)s->rst_cap * m.nq); s->rst_qvel_dev = dalloc<float>(s, (size_t)s->rst_cap * m.nv);
  }
  upload_async(s, s->rst_ids_dev, ids.data(), sizeof(int) * cnt);
  upload_async(s, s->rst_qpos_dev, qp.data(), sizeof(float) * qp.size());
  if (qvel) upload_async(s, s->rst_qvel_dev, qv.data(), sizeof(float) * qv.size());
#ifndef FB_EMU
  cudaStreamSynchronize(s->stream);       // the host vectors go out of scope below
#endif
  s->d.rst_ids = s->rst_ids_dev; s->d.rst_qpos = s->rst_qpos_dev; s->d.rst_qvel = s->rst_qvel_dev; s->d.rst_n = cnt; s->d.rst_has_qvel = qvel ? 1 : 0; s->d.rst_hold = hold;
  fb_launch<ShNone, Ph<ph_reset_scatter>>(s, K_MISC, 0, cnt);
  s->d.rst_n = 0;
  return 0;
}
int fb_reset(FbHandle s, const int32_t* env_ids, int n, const float* qpos, const float* qvel) {
  if (!s) return -1;
  int rc = do_reset(s, env_ids, n, qpos, qvel, 0);
  return rc != 0 ? rc : fb_forward(s);
}
int fb_reset_hold(FbHandle s, const int32_t* env_ids, int n, const float* qpos, const float* qvel) {
  if (!s || !env_ids || !qpos || n <= 0) return -1;
  int rc = do_reset(s, env_ids, n, qpos, qvel, 1);
  if (rc == 0) s->hold_pending = 1;
  return rc;
}

int fb_profile(FbHandle s, int enable) {
  if (!s) return -1;
  s->prof_on = enable ? 1 : 0;
  if (enable) { memset(s->prof_ms, 0, sizeof(s->prof_ms)); memset(s->prof_n, 0, sizeof(s->prof_n)); }
  return 0;
}
int fb_profile_read(FbHandle s, double* ms, long long* counts, int n) {
  if (!s || !ms || !counts) return -1;
#ifndef FB_EMU
  if (sync_stream(s) != 0) return -2;
  for (auto& ev : s->prof_events) { float t = 0; cudaEventElapsedTime(&t, ev.a, ev.b); s->prof_ms[ev.kind] += t; s->prof_n[ev.kind]++; cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  s->prof_events.clear();
#endif
  for (int k = 0; k < n && k < K_NKIND; k++) { ms[k] = s->prof_ms[k]; counts[k] = s->prof_n[k]; }
  return K_NKIND;
}
const char* fb_profile_name(int kind) { return (kind >= 0 && kind < K_NKIND) ? kKindNames[kind] : ""; }

int fb_obs_program(FbHandle s, const FbObsProgram* p) {
  if (!s || !p || p->n_items <= 0) return -1;
  if (sync_stream(s) != 0) return -2;
  const DevModel& m = s->m;
  int dim = 0; std::vector<int> offs;
  for (int i = 0; i < p->n_items; i++) {
    int k = p->kind[i], b = p->b[i];
    offs.push_back(dim);
    switch (k) {
      case FB_OBS_SENSOR_MEAN: case FB_OBS_SENSOR_NOW: case FB_OBS_ACT: case FB_OBS_QPOS: case FB_OBS_QVEL: case FB_OBS_TASK_TARGET: dim += b; break;
      case FB_OBS_SITES_EGO: case FB_OBS_REF_DISP: case FB_OBS_DOF_AXIS_EGO: dim += 3 * b; break;
      case FB_OBS_REF_QUAT: dim += 4 * b; break;
      case FB_OBS_ROOT_ZAXIS: case FB_OBS_SCALARS: case FB_OBS_SUBTREE_COM: dim += 3; break;
      case FB_OBS_ROOT_POSE: dim += 7; break;
      case FB_OBS_WORLD_CONTACT: dim += 1; break;
      default: s->err = "fb_obs_program: unknown item kind"; return -1;
    }
    if ((k == FB_OBS_REF_DISP || k == FB_OBS_REF_QUAT) && (!p->ref_qpos || p->ref_len <= 0)) { s->err = "fb_obs_program: reference table missing"; return -1; }
    if ((k == FB_OBS_QPOS || k == FB_OBS_QVEL || k == FB_OBS_SITES_EGO || k == FB_OBS_DOF_AXIS_EGO) && (p->a[i] < 0 || b < 0 || p->a[i] + b > p->n_list)) { s->err = "fb_obs_program: list range"; return -1; }
    if ((k == FB_OBS_SENSOR_MEAN || k == FB_OBS_SENSOR_NOW) && (p->a[i] < 0 || b < 0 || p->a[i] + b > m.nsensordata)) { s->err = "fb_obs_program: sensor range"; return -1; }
    if (k == FB_OBS_ACT && (p->a[i] < 0 || b < 0 || p->a[i] + b > m.na)) { s->err = "fb_obs_program: activation range"; return -1; }
    if (k == FB_OBS_SUBTREE_COM && (p->a[i] < 0 || p->a[i] >= m.nbody)) { s->err = "fb_obs_program: subtree_com body id"; return -1; }
    if ((k == FB_OBS_REF_DISP || k == FB_OBS_REF_QUAT) && b < 0) { s->err = "fb_obs_program: negative count"; return -1; }
  }
  for (int i = 0; i < p->n_items; i++) {          // entries of the index list, by the kind that reads them
    const int k = p->kind[i], lim = k == FB_OBS_QPOS ? m.nq : (k == FB_OBS_QVEL || k == FB_OBS_DOF_AXIS_EGO) ? m.nv : k == FB_OBS_SITES_EGO ? m.nsite : -1;
    if (lim < 0) continue;
    for (int j = 0; j < p->b[i]; j++) if (p->list[p->a[i] + j] < 0 || p->list[p->a[i] + j] >= lim) { s->err = "fb_obs_program: list entry out of range"; return -1; }
  }
  if (p->root_body <= 0 || p->root_body >= m.nbody) { s->err = "fb_obs_program: root body"; return -1; }
  // a replaced program's buffers are freed; the task program points into them (observation rows, per-env step inputs), so it goes too
  // and has to be uploaded again (fly_envs does: observation program, then task program)
  scope_free(s, s->task_scope); s->d.task = nullptr;
  scope_free(s, s->obs_scope);
  const size_t obs_mark = s->allocs.size();
  std::vector<int> kind(p->kind, p->kind + p->n_items), a(p->a, p->a + p->n_items), b(p->b, p->b + p->n_items), list(p->list, p->list + std::max(p->n_list, 0));
  s->d.op_kind = up(s, kind); s->d.op_a = up(s, a); s->d.op_b = up(s, b); s->d.op_off = up(s, offs); s->d.op_list = up(s, list);
  s->d.op_n = p->n_items; s->d.op_root_body = p->root_body; s->d.op_nsub = p->n_sub; s->d.op_ref_len = p->ref_len; s->d.op_ref_slot = 0;
  if (p->ref_qpos && p->ref_len > 0) { std::vector<float> r(p->ref_qpos, p->ref_qpos + (size_t)7 * p->ref_len); s->d.op_ref = up(s, r); } else s->d.op_ref = nullptr;
  s->d.tobs_dim = dim; s->d.tobs = dalloc<float>(s, (size_t)dim * s->d.Np);
  s->op_step_dev = dalloc<int>(s, s->d.Np); s->op_first_dev = (unsigned char*)dalloc<int>(s, s->d.Np);
  s->d.op_step = s->op_step_dev; s->d.op_first = s->op_first_dev;
  s->obs_scope.assign(s->allocs.begin() + obs_mark, s->allocs.end());
#ifndef FB_EMU
  cudaDeviceSynchronize();
#endif
  return dim;
}
int fb_ref_slots(FbHandle s, int slot_len) {
  if (!s || !s->d.tobs || slot_len <= 0) return -1;
  if (sync_stream(s) != 0) return -2;
  if (!s->ref_slots || s->ref_slot_len != slot_len) {          // (re)allocate: one [slot_len][7] table per env
    if (s->ref_slots) {
      for (auto& a : s->allocs) if (a == s->ref_slots) { a = nullptr; break; }
      dev_free(s->ref_slots);
    }
    s->ref_slots = dalloc<float>(s, (size_t)s->d.Np * 7 * slot_len); s->ref_slot_len = slot_len;
  }
  s->d.op_ref = s->ref_slots; s->d.op_ref_len = slot_len; s->d.op_ref_slot = 1;
  return 0;
}
int fb_ref_slot_write(FbHandle s, const int32_t* env_ids, int n, const float* rows) {
  if (!s || !s->ref_slots || !s->d.op_ref_slot || !env_ids || !rows || n < 0) return -1;
  const size_t per = (size_t)7 * s->ref_slot_len;
  for (int k = 0; k < n; k++) {
    if (env_ids[k] < 0 || env_ids[k] >= s->d.N) { s->err = "fb_ref_slot_write: env id out of range"; return -1; }
    upload_async(s, s->ref_slots + per * env_ids[k], rows + per * k, sizeof(float) * per);
  }
  return sync_stream(s);                                        // the caller's rows may be reused right away
}
int fb_task_program(FbHandle s, const FbTaskProgram* p) {
  if (!s || !p) return -1;
  if (!s->d.tobs || !s->act_map_dev) { s->err = "fb_task_program: call fb_set_action_map and fb_obs_program first"; return -1; }
  if (s->d.op_ref_slot) { s->err = "fb_task_program: per-env reference slots keep the host-side task code"; return -1; }
  if (p->ref_len <= 0 || !p->ref_qpos || !p->ref_qvel || !p->reset_qpos) { s->err = "fb_task_program: reference / reset tables missing"; return -1; }
  if (p->kind == 2 && (!s->bank_dev || !s->hfield_dev || s->bank_n <= 0)) { s->err = "fb_task_program: kind 2 needs fb_hfield_collision / fb_eye_program and fb_hfield_bank first"; return -1; }
  if (p->kind >= 1 && (p->n_wing <= 0 || p->n_freq <= 0 || p->tab_len <= 0 || !p->wb_traj || !p->wb_phase || !p->wb_phase_mod || !p->wb_freqs || !p->wb_len)) { s->err = "fb_task_program: wing-beat tables missing"; return -1; }
  if (sync_stream(s) != 0) return -2;
  const DevModel& m = s->m; const int Np = s->d.Np;
  scope_free(s, s->task_scope); s->d.task = nullptr;
  const size_t task_mark = s->allocs.size();
  DevTask t; memset(&t, 0, sizeof(t));
  t.kind = p->kind; t.root_qadr = p->root_qadr; t.root_vadr = p->root_vadr; t.ghost_qadr = p->ghost_qadr; t.ghost_vadr = p->ghost_vadr;
  t.root_body = s->d.op_root_body; t.user_col = p->user_col;
  for (int i = 0; i < 3; i++) { t.ghost_offset[i] = p->ghost_offset[i]; t.com_offset[i] = p->com_offset[i]; }
  t.dt = p->control_timestep; t.time_limit = p->time_limit; t.term_com = p->terminal_com_dist; t.term_linvel = p->terminal_linvel;
  t.term_angvel = p->terminal_angvel; t.term_qacc = p->terminal_qacc; t.term_height = p->terminal_height;
  t.velocimeter_adr = p->velocimeter_adr; t.gyro_adr = p->gyro_adr; t.com_body = p->com_body; t.episode_steps = p->episode_steps;
  t.ref_len = p->ref_len; t.obs_refdisp_off = p->obs_refdisp_off; t.obs_refquat_off = p->obs_refquat_off;
  auto upf32 = [&](const float* src, size_t n) { std::vector<float> v(src, src + n); return up(s, v); };
  auto upi32 = [&](const int32_t* src, size_t n) { std::vector<int> v(src, src + n); return up(s, v); };
  t.reset_qpos = upf32(p->reset_qpos, m.nq); t.ref_qpos = upf32(p->ref_qpos, (size_t)7 * p->ref_len); t.ref_qvel = upf32(p->ref_qvel, (size_t)6 * p->ref_len);
  t.n_noise = p->noise_qadr ? p->n_noise : 0; t.noise_qadr = t.n_noise ? upi32(p->noise_qadr, t.n_noise) : nullptr; t.noise_amp = p->noise_amp; t.seed = p->seed;
  if (p->kind >= 1) {
    t.n_wing = p->n_wing; t.wing_qadr = upi32(p->wing_qadr, p->n_wing); t.wing_vadr = upi32(p->wing_vadr, p->n_wing); t.wing_ctrl = upi32(p->wing_ctrl, p->n_wing);
    t.n_freq = p->n_freq; t.tab_len = p->tab_len; const size_t nt = (size_t)p->n_freq * p->tab_len;
    t.wb_traj = upf32(p->wb_traj, nt * p->n_wing); t.wb_phase = upf32(p->wb_phase, nt); t.wb_phase_mod = upf32(p->wb_phase_mod, nt);
    t.wb_freqs = upf32(p->wb_freqs, p->n_freq); t.wb_len = upi32(p->wb_len, p->n_freq);
    t.wb_base_freq = p->wb_base_freq; t.wb_rel_range = p->wb_rel_range; t.wb_rate = p->wb_rate;
  }
  t.step = dalloc<int>(s, Np); t.needs_reset = dalloc<int>(s, Np); t.resetting = dalloc<int>(s, Np); t.episode = dalloc<int>(s, Np);
  t.wb_idx = dalloc<int>(s, Np); t.wb_pos = dalloc<int>(s, Np); t.has_uniform = dalloc<int>(s, Np);
  t.uniform = dalloc<float>(s, (size_t)8 * Np); t.wb_freq = dalloc<float>(s, Np); t.out = dalloc<float>(s, (size_t)4 * Np);
  if (p->kind == 2) {
    for (int i = 0; i < 2; i++) { t.th_rng[i] = p->target_height_range[i]; t.ts_rng[i] = p->target_speed_range[i]; t.x_rng[i] = p->init_x_range[i]; t.y_rng[i] = p->init_y_range[i]; }
    for (int i = 0; i < 4; i++) t.hover_quat[i] = p->hover_quat[i];
    for (int i = 0; i < 3; i++) t.target_zaxis[i] = p->target_zaxis[i];
    t.fatal = p->floor_contacts_fatal;
    t.n_bank = s->bank_n; t.hf_nrow = s->hf_nrow; t.hf_ncol = s->hf_ncol; t.hf_ncm = s->hf_nbr * s->hf_nbc; t.hf_half = s->hf.size[0]; t.hf_zoff = 0.0f;
    t.bank = s->bank_dev; t.bank_hmax = s->bank_hmax_dev; t.bank_cmax = s->bank_cmax_dev; t.hf_data = s->hfield_dev; t.hf_hmax = s->hmax_dev; t.hf_cmax = s->cmax_dev;
    t.target = dalloc<float>(s, (size_t)2 * Np); t.pick = dalloc<int>(s, Np);
    t.trench_cap = 0; t.trench_x = t.trench_y = nullptr; t.trench_len = nullptr;
    if (p->trench_cap > 0) {
      if (!p->trench_y || !p->trench_x || !p->trench_len) { s->err = "fb_task_program: trench centre lines missing"; return -1; }
      for (int k = 0; k < s->bank_n; k++) if (p->trench_len[k] < 2 || p->trench_len[k] > p->trench_cap || !(p->trench_x[2 * k + 1] > p->trench_x[2 * k])) { s->err = "fb_task_program: bad trench centre line"; return -1; }
      float* ty = dalloc<float>(s, (size_t)p->trench_cap * s->bank_n); h2d(ty, p->trench_y, sizeof(float) * (size_t)p->trench_cap * s->bank_n);
      float* tx = dalloc<float>(s, (size_t)2 * s->bank_n); h2d(tx, p->trench_x, sizeof(float) * 2 * s->bank_n);
      int* tl = dalloc<int>(s, s->bank_n); h2d(tl, p->trench_len, sizeof(int) * s->bank_n);
      t.trench_cap = p->trench_cap; t.trench_x = tx; t.trench_y = ty; t.trench_len = tl;
    }
  }
  t.op_step = s->op_step_dev; t.op_first = s->op_first_dev;
  DevTask* dev = (DevTask*)dalloc<unsigned char>(s, sizeof(DevTask));
  h2d(dev, &t, sizeof(t));
  s->task_host = t; s->d.task = dev;
  s->task_scope.assign(s->allocs.begin() + task_mark, s->allocs.end());
  return fb_task_reset_all(s);
}
int fb_task_reset_all(FbHandle s) {
  if (!s || !s->d.task) return -1;
  if (sync_stream(s) != 0) return -2;
  std::vector<int> ones((size_t)s->d.Np, 1);
  h2d(s->task_host.needs_reset, ones.data(), sizeof(int) * ones.size());
  return 0;
}
int fb_task_set_reset_noise(FbHandle s, float amp) {
  if (!s || !s->d.task || !(amp >= 0)) return -1;
  if (sync_stream(s) != 0) return -2;
  s->task_host.noise_amp = amp;
  h2d((void*)s->d.task, &s->task_host, sizeof(DevTask));          // same pointers, new scalar
  return 0;
}
int fb_task_request_reset(FbHandle s, const int32_t* env_ids, int n) {
  if (!s || !s->d.task || (n > 0 && !env_ids) || n < 0) return -1;
  if (n == 0) return 0;
  if (sync_stream(s) != 0) return -2;
  const int one = 1;
  for (int k = 0; k < n; k++) {
    if (env_ids[k] < 0 || env_ids[k] >= s->d.N) { s->err = "fb_task_request_reset: env id out of range"; return -1; }
    h2d(s->task_host.needs_reset + env_ids[k], &one, sizeof(int));
  }
  return 0;
}
int fb_task_episodes(FbHandle s, int32_t* dst) {
  if (!s || !s->d.task || !dst) return -1;
  if (sync_stream(s) != 0) return -2;
  d2h(dst, s->task_host.episode, sizeof(int) * (size_t)s->d.N);
  return 0;
}
int fb_task_uniforms(FbHandle s, const int32_t* env_ids, int n, const float* u) {
  if (!s || !s->d.task || !env_ids || !u || n < 0) return -1;
  if (sync_stream(s) != 0) return -2;
  const int one = 1;
  for (int k = 0; k < n; k++) {
    if (env_ids[k] < 0 || env_ids[k] >= s->d.N) { s->err = "fb_task_uniforms: env id out of range"; return -1; }
    h2d(s->task_host.uniform + (size_t)8 * env_ids[k], u + k, sizeof(float)); h2d(s->task_host.has_uniform + env_ids[k], &one, sizeof(int));
  }
  return 0;
}
int fb_task_uniform_rows(FbHandle s, const int32_t* env_ids, int n, const float* u) {
  if (!s || !s->d.task || !env_ids || !u || n < 0) return -1;
  if (sync_stream(s) != 0) return -2;
  const int one = 1;
  for (int k = 0; k < n; k++) {
    if (env_ids[k] < 0 || env_ids[k] >= s->d.N) { s->err = "fb_task_uniform_rows: env id out of range"; return -1; }
    h2d(s->task_host.uniform + (size_t)8 * env_ids[k], u + (size_t)8 * k, 8 * sizeof(float)); h2d(s->task_host.has_uniform + env_ids[k], &one, sizeof(int));
  }
  return 0;
}
int fb_hfield_bank(FbHandle s, int n_terrain, const float* heights) {
  if (!s || !s->hfield_dev || n_terrain <= 0 || !heights) { if (s) s->err = "fb_hfield_bank: set up the heightfield grid first (fb_hfield_collision / fb_eye_program)"; return -1; }
  if (sync_stream(s) != 0) return -2;
  const size_t cells = (size_t)s->hf_nrow * s->hf_ncol, ncm = (size_t)s->hf_nbr * s->hf_nbc;
  std::vector<float> hmax(n_terrain), cm(ncm * n_terrain);
  for (int k = 0; k < n_terrain; k++) {
    const float* hk = heights + cells * k;
    float mx = hk[0]; for (size_t i = 1; i < cells; i++) mx = std::max(mx, hk[i]);
    hmax[k] = mx;
    for (int br = 0; br < s->hf_nbr; br++) for (int bc = 0; bc < s->hf_nbc; bc++) {
      float v = -3.0e38f;
      for (int iy = br * FB_EYE_BLOCK; iy <= std::min((br + 1) * FB_EYE_BLOCK, s->hf_nrow - 1); iy++)
        for (int ix = bc * FB_EYE_BLOCK; ix <= std::min((bc + 1) * FB_EYE_BLOCK, s->hf_ncol - 1); ix++) v = std::max(v, hk[(size_t)iy * s->hf_ncol + ix]);
      cm[ncm * k + (size_t)br * s->hf_nbc + bc] = v;
    }
  }
  if (s->bank_dev) {      // a replaced bank is freed; a vision task program pointing at it has to be uploaded again
    std::vector<void*> old = {s->bank_dev, s->bank_hmax_dev, s->bank_cmax_dev};
    scope_free(s, old);
    if (s->d.task && s->task_host.kind == 2) { scope_free(s, s->task_scope); s->d.task = nullptr; }
  }
  s->bank_dev = dalloc<float>(s, cells * n_terrain); s->bank_hmax_dev = dalloc<float>(s, n_terrain); s->bank_cmax_dev = dalloc<float>(s, ncm * n_terrain);
  if (!s->bank_dev) { s->err = "out of device memory (terrain bank)"; return -4; }
  h2d(s->bank_dev, heights, sizeof(float) * cells * n_terrain); h2d(s->bank_hmax_dev, hmax.data(), sizeof(float) * n_terrain); h2d(s->bank_cmax_dev, cm.data(), sizeof(float) * cm.size());
  s->bank_n = n_terrain;
  return 0;
}
int fb_task_step(FbHandle s, const float* action, int is_device, int n_substeps) {
  if (!s || !s->d.task || !action || n_substeps <= 0) return -1;
#ifndef FB_EMU
  cudaSetDevice(s->device);
#endif
  fb_launch<ShNone, Ph<ph_task_reset>, Ph<ph_task_reset2>>(s, K_MISC, 0, s->d.N);      // envs whose last step was LAST
  int rc = fb_set_ctrl(s, action, is_device);                                           // action -> ctrl (NaN -> 0), staged rows stay readable
  if (rc != 0) return rc;
  fb_launch<ShNone, Ph<ph_task_before>, Ph<ph_task_commit>>(s, K_MISC, 0, s->d.N);
  s->hold_pending = 1;                                                                  // freshly reset envs are held through this step
  rc = fb_step(s, n_substeps);
  if (rc != 0) return rc;
  fb_launch<ShNone, Ph<ph_pack>, Ph<ph_task_after>>(s, K_PACK, 0, s->d.N);
  return 0;
}
int fb_task_ptrs(FbHandle s, void** obs_dev, int* obs_dim, void** out_dev) {
  if (!s || !s->d.task) return -1;
  if (obs_dev) *obs_dev = s->d.tobs;
  if (obs_dim) *obs_dim = s->d.tobs_dim;
  if (out_dev) *out_dev = s->task_host.out;
  return 0;
}
int fb_task_read(FbHandle s, float* obs_host, float* out_host) {
  if (!s || !s->d.task) return -1;
#ifndef FB_EMU
  if (obs_host) FB_CUDA_OK(cudaMemcpyAsync(obs_host, s->d.tobs, sizeof(float) * (size_t)s->d.N * s->d.tobs_dim, cudaMemcpyDeviceToHost, s->stream));
  if (out_host) FB_CUDA_OK(cudaMemcpyAsync(out_host, s->task_host.out, sizeof(float) * (size_t)4 * s->d.N, cudaMemcpyDeviceToHost, s->stream));
  return sync_stream(s);
#else
  if (obs_host) memcpy(obs_host, s->d.tobs, sizeof(float) * (size_t)s->d.N * s->d.tobs_dim);
  if (out_host) memcpy(out_host, s->task_host.out, sizeof(float) * (size_t)4 * s->d.N);
  return 0;
#endif
}
#ifndef FB_EMU
__global__ void __launch_bounds__(128) fb_render_kernel(DevData d, DevEye p) {
  const int e = blockIdx.x / p.n_cam, cam = blockIdx.x % p.n_cam;
  for (int px = threadIdx.x; px < p.size * p.size; px += blockDim.x) eye_pixel(d, p, e, cam, px / p.size, px % p.size);
}
#endif
// per-env heightfield buffer [Np][nrow * ncol] + highest point per env, shared by the collision and the eye cameras
static int ensure_hfield(FbSim* s, int nrow, int ncol) {
  if (s->hfield_dev) {
    if (s->hf_nrow != nrow || s->hf_ncol != ncol) { s->err = "heightfield grid differs from the one already allocated"; return -1; }
    return 0;
  }
  s->hfield_dev = dalloc<float>(s, (size_t)nrow * ncol * s->d.Np); s->hmax_dev = dalloc<float>(s, s->d.Np);
  s->hf_nrow = nrow; s->hf_ncol = ncol;
  // block maxima for the eye ray marcher (fb_render.h: FB_EYE_BLOCK)
  s->hf_nbr = (nrow - 1 + FB_EYE_BLOCK - 1) / FB_EYE_BLOCK; s->hf_nbc = (ncol - 1 + FB_EYE_BLOCK - 1) / FB_EYE_BLOCK;
  s->cmax_dev = dalloc<float>(s, (size_t)s->hf_nbr * s->hf_nbc * s->d.Np);
  return 0;
}
int fb_hfield_collision(FbHandle s, int geom, const float* size, int nrow, int ncol, const int32_t* pair_geom, int npair) {
  if (!s || !size || !pair_geom || npair <= 0 || npair > FB_MAXCAND || nrow < 2 || ncol < 2) return -1;
  if (geom < 0 || geom >= s->m.ngeom) { s->err = "fb_hfield_collision: geom id out of range"; return -1; }
  for (int k = 0; k < npair; k++) if (pair_geom[k] < 0 || pair_geom[k] >= s->m.ngeom) { s->err = "fb_hfield_collision: pair geom out of range"; return -1; }
  if (s->fuse != 0) { s->err = "fb_hfield_collision: needs the one-kernel-per-stage launch sequence (FB_FUSE=0, FB_SPLIT=1)"; return -1; }
#ifndef FB_EMU
  if (s->split > 1) { s->err = "fb_hfield_collision: needs the one-kernel-per-stage launch sequence (FB_FUSE=0, FB_SPLIT=1)"; return -1; }
#endif
  if (sync_stream(s) != 0) return -2;
  if (ensure_hfield(s, nrow, ncol) != 0) return -1;
  DevHf& h = s->hf; h.geom = geom; h.nrow = nrow; h.ncol = ncol; h.npair = npair;
  for (int k = 0; k < 4; k++) h.size[k] = size[k];
  std::vector<int> pg(pair_geom, pair_geom + npair); h.pair_geom = up(s, pg);
  h.data = s->hfield_dev; h.hmax = s->hmax_dev;
  s->hf_on = true;
#ifndef FB_EMU
  for (auto& g : s->graph) { if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; } g.seen = false; }      // the launch sequence changed
#endif
  return 0;
}
int fb_eye_program(FbHandle s, const FbEyeProgram* p) {
  if (!s || !p) return -1;
  if (p->n_cam < 1 || p->n_cam > FB_MAXCAM || p->size < 1 || p->size > 512 || !(p->fovy_deg > 0 && p->fovy_deg < 180) || p->nrow < 0 || (p->nrow > 0 && (p->nrow < 2 || p->ncol < 2 || !(p->half_size > 0)))) { s->err = "fb_eye_program: bad camera / grid parameters"; return -1; }
  for (int c = 0; c < p->n_cam; c++) if (p->body[c] <= 0 || p->body[c] >= s->m.nbody) { s->err = "fb_eye_program: camera body out of range"; return -1; }
  if (sync_stream(s) != 0) return -2;
  DevEye& y = s->eye; memset(&y, 0, sizeof(y));
  y.n_cam = p->n_cam; y.size = p->size; y.nrow = p->nrow; y.ncol = p->nrow > 0 ? p->ncol : 0;
  for (int c = 0; c < p->n_cam; c++) { y.body[c] = p->body[c]; for (int k = 0; k < 3; k++) y.pos[c][k] = p->pos[c][k]; for (int k = 0; k < 4; k++) y.quat[c][k] = p->quat[c][k]; }
  y.tan_half = tanf(0.5f * p->fovy_deg * 3.14159265358979f / 180.0f); y.half_size = p->half_size; y.z_offset = p->z_offset; y.zfar = p->zfar > 0 ? p->zfar : 50.0f;
  for (int k = 0; k < 3; k++) { y.sky_top[k] = p->sky_top[k]; y.sky_horizon[k] = p->sky_horizon[k]; y.ground[k] = p->ground[k]; }
  y.ambient = p->ambient; y.diffuse = p->diffuse;
  if (y.nrow > 0 && ensure_hfield(s, y.nrow, y.ncol) != 0) return -1;
  s->eye_bytes = (size_t)y.n_cam * y.size * y.size * 3;
  s->eye_out = dalloc<unsigned char>(s, s->eye_bytes * s->d.Np);
  y.hfield = s->hfield_dev; y.hmax = s->hmax_dev; y.out = s->eye_out;
  y.cmax = (y.nrow > 0 && !getenv("FB_EYE_NO_SKIP")) ? s->cmax_dev : nullptr; y.nbr = s->hf_nbr; y.nbc = s->hf_nbc;
  return 0;
}
int fb_hfield_write(FbHandle s, const int32_t* env_ids, int n, const float* heights) {
  if (!s || !s->hfield_dev || !env_ids || !heights || n < 0) return -1;
  const size_t cells = (size_t)s->hf_nrow * s->hf_ncol;
  for (int k = 0; k < n; k++) {
    if (env_ids[k] < 0 || env_ids[k] >= s->d.N) { s->err = "fb_hfield_write: env id out of range"; return -1; }
    float mx = heights[cells * k]; for (size_t i = 1; i < cells; i++) mx = std::max(mx, heights[cells * k + i]);
    upload_async(s, s->hfield_dev + cells * env_ids[k], heights + cells * k, sizeof(float) * cells);
    // highest grid point of every block of FB_EYE_BLOCK x FB_EYE_BLOCK cells, border points included
    std::vector<float> cm((size_t)s->hf_nbr * s->hf_nbc);
    const float* hk = heights + cells * k;
    for (int br = 0; br < s->hf_nbr; br++) for (int bc = 0; bc < s->hf_nbc; bc++) {
      float v = -3.0e38f;
      for (int iy = br * FB_EYE_BLOCK; iy <= std::min((br + 1) * FB_EYE_BLOCK, s->hf_nrow - 1); iy++)
        for (int ix = bc * FB_EYE_BLOCK; ix <= std::min((bc + 1) * FB_EYE_BLOCK, s->hf_ncol - 1); ix++) v = std::max(v, hk[(size_t)iy * s->hf_ncol + ix]);
      cm[(size_t)br * s->hf_nbc + bc] = v;
    }
    if (sync_stream(s) != 0) return -2;
    h2d(s->cmax_dev + cm.size() * env_ids[k], cm.data(), sizeof(float) * cm.size());
    h2d(s->hmax_dev + env_ids[k], &mx, sizeof(float));
  }
  return 0;
}
int fb_render_eyes(FbHandle s) {
  if (!s || !s->eye_out) return -1;
#ifndef FB_EMU
  cudaSetDevice(s->device);
  fb_render_kernel<<<s->d.N * s->eye.n_cam, 128, 0, s->stream>>>(s->d, s->eye);
  if (cudaGetLastError() != cudaSuccess) { s->err = "fb_render_eyes: launch failed"; return -2; }
#else
  for (int e = 0; e < s->d.N; e++) for (int c = 0; c < s->eye.n_cam; c++) for (int px = 0; px < s->eye.size * s->eye.size; px++) eye_pixel(s->d, s->eye, e, c, px / s->eye.size, px % s->eye.size);
#endif
  s->launches++;
  return 0;
}
int fb_eyes_ptr(FbHandle s, void** dev_ptr, int* bytes_per_env) {
  if (!s || !s->eye_out) return -1;
  if (dev_ptr) *dev_ptr = s->eye_out;
  if (bytes_per_env) *bytes_per_env = (int)s->eye_bytes;
  return 0;
}
int fb_eyes_read(FbHandle s, uint8_t* host_dst) {
  if (!s || !s->eye_out || !host_dst) return -1;
#ifndef FB_EMU
  FB_CUDA_OK(cudaMemcpyAsync(host_dst, s->eye_out, s->eye_bytes * s->d.N, cudaMemcpyDeviceToHost, s->stream));
  return sync_stream(s);
#else
  memcpy(host_dst, s->eye_out, s->eye_bytes * s->d.N);
  return 0;
#endif
}
int fb_task_inputs(FbHandle s, const int32_t* step_idx, const uint8_t* first) {
  if (!s || !s->d.tobs || !step_idx || !first) return -1;
  upload_async(s, s->op_step_dev, step_idx, sizeof(int) * s->d.N);
  upload_async(s, s->op_first_dev, first, s->d.N);
  return 0;
}
int fb_read_task_obs(FbHandle s, float* host_dst) {
  if (!s || !host_dst || !s->d.tobs) return -1;
#ifndef FB_EMU
  FB_CUDA_OK(cudaMemcpyAsync(host_dst, s->d.tobs, sizeof(float) * (size_t)s->d.N * s->d.tobs_dim, cudaMemcpyDeviceToHost, s->stream));
  return sync_stream(s);
#else
  memcpy(host_dst, s->d.tobs, sizeof(float) * (size_t)s->d.N * s->d.tobs_dim);
  return 0;
#endif
}
int fb_pack_obs(FbHandle s) {
  if (!s) return -1;
#ifndef FB_EMU
  cudaSetDevice(s->device);
#endif
  fb_launch<ShNone, Ph<ph_pack>>(s, K_PACK);
  return 0;
}
int fb_read_obs(FbHandle s, float* host_dst) {
  if (!s || !host_dst) return -1;
#ifndef FB_EMU
  FB_CUDA_OK(cudaMemcpyAsync(host_dst, s->d.obs, sizeof(float) * (size_t)s->d.N * s->d.obs_dim, cudaMemcpyDeviceToHost, s->stream));
  return sync_stream(s);
#else
  memcpy(host_dst, s->d.obs, sizeof(float) * (size_t)s->d.N * s->d.obs_dim);
  return 0;
#endif
}
int fb_obs_ptr(FbHandle s, void** dev_ptr, int* floats_per_env) {
  if (!s || !dev_ptr || !floats_per_env) return -1;
  *dev_ptr = s->d.obs; *floats_per_env = s->d.obs_dim;
  return 0;
}

}  // extern "C"
