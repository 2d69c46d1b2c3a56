// csrc/api.cu -- extern "C" entry points of the compute path and the step orchestration.
//
// Orchestration replaces GetIntermidiateVelocityContactResolved (self_collision_cloth.cpp:355-456) and the demo
// StepTime functions (self_contact_sparse/main.cpp:213-263, internal_cloth_sparse/main.cpp:205-237).  The loop
// control of the reference depends on the number of contact elements found in each pass (:416, :449); those counts
// are read back once per pass (4 bytes per scene) -- the only device->host traffic of a step.
#include <cstring>

#include "ctx.h"

namespace cg {


// the active flags of up to 64 scenes travel as a kernel argument: no staging buffer, no host synchronisation
__global__ void k_set_active_mask(int B, int* cnt, unsigned long long mask) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < B) cnt[s * NCNT + CNT_ACTIVE] = (int)((mask >> s) & 1ULL);
}

int set_active(Ctx* c, const int* active) {
  for (int s = 0; s < c->B; s++) c->h_cnt[s * NCNT + CNT_ACTIVE] = active ? (active[s] != 0) : 1;
  if (c->B <= 64) {
    unsigned long long mask = 0;
    for (int s = 0; s < c->B; s++) mask |= (unsigned long long)(c->h_cnt[s * NCNT + CNT_ACTIVE] != 0) << s;
    k_set_active_mask<<<1, 64, 0, KS(c, "k_set_active_mask")>>>(c->B, c->cnt, mask);
    CG_LAUNCH_CHECK(c);
    return CLOTHGPU_OK;
  }
  CG_CUDA(cudaMemcpy2DAsync(c->cnt + CNT_ACTIVE, sizeof(int) * NCNT, c->h_cnt + CNT_ACTIVE, sizeof(int) * NCNT,
                            sizeof(int), c->B, cudaMemcpyHostToDevice, c->stream));
  // h_cnt is reused by the next counter fetch: the upload must have left the host buffer before that
  CG_CUDA(cudaStreamSynchronize(c->stream));
  return CLOTHGPU_OK;
}

static int copy_vec(Ctx* c, const double* src, double* dst) {
  CG_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * (size_t)c->B * c->ndof, cudaMemcpyDeviceToDevice, c->stream));
  return CLOTHGPU_OK;
}

int run_resolve(Ctx* c, const clothgpu_params& p, int* applied, int* stats, bool prox_prelaunched) {
  if (c->nnode == 0) { set_error("context has no BVH"); return CLOTHGPU_ERR_STATE; }
  const int B = c->B;
  std::vector<int> act(B, 1), n(B, 0), app(B, 0);
  if (stats) memset(stats, 0, sizeof(int) * CLOTHGPU_NSTATS * B);
  const double* x = c->xyz_prev;
  double* v = c->uvw;
  const double dt = p.dt, delta = p.contact_clearance;
  // proximity pass, :369-395.  With prox_prelaunched the refit, traversal and narrow phase of this pass were issued
  // on the second stream before the implicit solve (they only read the positions at the start of the step) and the
  // main stream has already joined it.
  if (!prox_prelaunched) {
    CG_CHECK(set_active(c, nullptr));
    phase_begin(c, 3);
    CG_CHECK(launch_refit(c, x, nullptr, delta, 0));
    phase_end(c, 3);
  }
  if (delta * 0.5 > 0) {  // CAABB3D::AddPoint ignores eps <= 0: all boxes stay empty and nothing intersects
    CG_CHECK(run_contacts(c, x, nullptr, dt, delta, 0, n.data(), prox_prelaunched));
  }
  for (int s = 0; s < B; s++) {
    app[s] = n[s] > 0;
    if (stats) stats[s * CLOTHGPU_NSTATS + 1] = n[s];
  }
  // (with a skipped detection the reference applies the proximity impulse to an empty set: the element list of an
  // earlier pass must not be picked up)
  if (delta * 0.5 > 0) {
    phase_begin(c, 6);
    CG_CHECK(launch_impulse(c, x, v, delta, p.stiff_contact, dt, p.mass_point, 0));
    phase_end(c, 6);
  }
  // CCD impulse passes, :398-425
  int nactive = B;
  for (int itr = 0; itr < 5 && nactive > 0; itr++) {
    // from the second pass on only the velocities of the vertices the impulses touched have changed: incremental
    const bool inc = c->incremental_ccd && itr > 0;
    phase_begin(c, 3);
    CG_CHECK(launch_refit(c, x, v, dt, 1, inc));
    phase_end(c, 3);
    CG_CHECK(run_contacts(c, x, v, dt, delta, 1, n.data(), false, inc));
    nactive = 0;
    for (int s = 0; s < B; s++) {
      if (!act[s]) continue;
      if (stats) { stats[s * CLOTHGPU_NSTATS + 8 + itr] = n[s]; stats[s * CLOTHGPU_NSTATS + 2] = itr + 1; }
      if (n[s] == 0) act[s] = 0;  // :416 early return: this scene skips the impact-zone loop
      else { app[s] = 1; nactive++; }
    }
    if (nactive == 0) break;
    CG_CHECK(set_active(c, act.data()));
    phase_begin(c, 6);
    CG_CHECK(launch_impulse(c, x, v, delta, p.stiff_contact, dt, p.mass_point, 1));
    phase_end(c, 6);
  }
  // rigid impact zones, :426-455
  if (nactive > 0) {
    CG_CHECK(copy_vec(c, v, c->uvw0));
    CG_CHECK(launch_riz_reset(c));
    for (int itr = 0; itr < 100 && nactive > 0; itr++) {
      phase_begin(c, 3);
      CG_CHECK(launch_refit(c, x, v, dt, 1, c->incremental_ccd));
      phase_end(c, 3);
      CG_CHECK(run_contacts(c, x, v, dt, delta, 1, n.data(), false, c->incremental_ccd));
      nactive = 0;
      for (int s = 0; s < B; s++) {
        if (!act[s]) continue;
        if (stats) {
          stats[s * CLOTHGPU_NSTATS + 16 + itr] = n[s];
          stats[s * CLOTHGPU_NSTATS + 3] = itr + 1;
          stats[s * CLOTHGPU_NSTATS + 5] = itr == 0 ? 0 : c->h_cnt[s * NCNT + CNT_RIZ_NODES];
          stats[s * CLOTHGPU_NSTATS + 128 + itr] = stats[s * CLOTHGPU_NSTATS + 5];
          if (n[s] == 0) stats[s * CLOTHGPU_NSTATS + 6] = 1;
        }
        if (n[s] == 0) act[s] = 0;
        else nactive++;
      }
      if (nactive == 0) break;
      CG_CHECK(set_active(c, act.data()));
      phase_begin(c, 7);
      CG_CHECK(launch_riz_make(c));
      CG_CHECK(launch_riz_apply(c, x, v));
      phase_end(c, 7);
    }
  }
  for (int s = 0; s < B; s++) {
    if (applied) applied[s] = app[s];
    if (stats) {
      stats[s * CLOTHGPU_NSTATS + 4] = app[s];
      stats[s * CLOTHGPU_NSTATS + 7] = c->h_cnt[s * NCNT + CNT_IMP_SLOW];  // as of the last counter fetch
      stats[s * CLOTHGPU_NSTATS + 14] = c->h_cnt[s * NCNT + CNT_RANK_BAD];
      stats[s * CLOTHGPU_NSTATS + 15] = c->h_cnt[s * NCNT + CNT_RANK_BAD + 1];
    }
  }
  return CLOTHGPU_OK;
}

__global__ void k_reintegrate(int n, double dt, const double* __restrict__ x1, const double* __restrict__ v,
                              double* __restrict__ x, const int* __restrict__ cnt) {
  if (!cnt[blockIdx.y * NCNT + CNT_ACTIVE]) return;
  const size_t off = (size_t)blockIdx.y * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    x[off + i] = x1[off + i] + v[off + i] * dt;
}

}  // namespace cg

using namespace cg;

#define REQUIRE_CTX(c)                                  \
  do {                                                  \
    if (!(c)) { set_error("null context"); return CLOTHGPU_ERR_ARG; } \
    cudaSetDevice((c)->device);                         \
  } while (0)
#define REQUIRE_BVH(c)                                                             \
  do {                                                                             \
    if ((c)->nnode == 0) { set_error("context was created without a BVH"); return CLOTHGPU_ERR_STATE; } \
  } while (0)

static int check_params(const clothgpu_params* p) {
  if (!p) { set_error("params is null"); return CLOTHGPU_ERR_ARG; }
  if (!(p->dt > 0) || !(p->mass_point > 0)) { set_error("dt and mass_point must be positive"); return CLOTHGPU_ERR_ARG; }
  if (p->collider_kind < 0 || p->collider_kind > CLOTHGPU_COLLIDER_TABLE) { set_error("unknown collider kind"); return CLOTHGPU_ERR_ARG; }
  return CLOTHGPU_OK;
}
static double* need_vec(clothgpu_ctx* c, int buf) {
  double* v = vec_buf(c, buf);
  if (!v) set_error("not a vector buffer id");
  return v;
}

extern "C" {

int clothgpu_assemble(clothgpu_ctx* c, const clothgpu_params* p) {
  REQUIRE_CTX(c);
  CG_CHECK(check_params(p));
  CG_CHECK(launch_assemble(c, *p, true));
  return clothgpu_synchronize(c);
}

int clothgpu_matvec(clothgpu_ctx* c, double alpha, int buf_x, double beta, int buf_y) {
  REQUIRE_CTX(c);
  double* x = need_vec(c, buf_x);
  double* y = need_vec(c, buf_y);
  if (!x || !y || x == y) { set_error("matvec needs two distinct vector buffers"); return CLOTHGPU_ERR_ARG; }
  CG_CHECK(launch_matvec(c, alpha, x, beta, y));
  return clothgpu_synchronize(c);
}

int clothgpu_dot(clothgpu_ctx* c, int buf_a, int buf_b, double* out) {
  REQUIRE_CTX(c);
  double* a = need_vec(c, buf_a);
  double* b = need_vec(c, buf_b);
  if (!a || !b || !out) { set_error("bad argument"); return CLOTHGPU_ERR_ARG; }
  CG_CHECK(launch_dot(c, a, b, SC_DOT));
  CG_CUDA(cudaMemcpyAsync(c->h_scal, c->scal, sizeof(double) * NSCAL * c->B, cudaMemcpyDeviceToHost, c->stream));
  CG_CUDA(cudaStreamSynchronize(c->stream));
  for (int s = 0; s < c->B; s++) out[s] = c->h_scal[s * NSCAL + SC_DOT];
  return CLOTHGPU_OK;
}

int clothgpu_axpy(clothgpu_ctx* c, double a, int buf_x, int buf_y) {
  REQUIRE_CTX(c);
  double* x = need_vec(c, buf_x);
  double* y = need_vec(c, buf_y);
  if (!x || !y || x == y) { set_error("axpy needs two distinct vector buffers"); return CLOTHGPU_ERR_ARG; }
  CG_CHECK(launch_axpy(c, a, x, y));
  return clothgpu_synchronize(c);
}

int clothgpu_ilu_setvalue(clothgpu_ctx* c) {
  REQUIRE_CTX(c);
  CG_CHECK(launch_ilu_setvalue(c));
  return clothgpu_synchronize(c);
}
int clothgpu_ilu_factor(clothgpu_ctx* c) {
  REQUIRE_CTX(c);
  CG_CHECK(launch_ilu_factor(c));
  return clothgpu_synchronize(c);
}
int clothgpu_ilu_solve(clothgpu_ctx* c, int buf_vec) {
  REQUIRE_CTX(c);
  double* v = need_vec(c, buf_vec);
  if (!v) return CLOTHGPU_ERR_ARG;
  CG_CHECK(launch_ilu_solve(c, v));
  return clothgpu_synchronize(c);
}

int clothgpu_pcg(clothgpu_ctx* c, double* conv_ratio, int* iteration) {
  REQUIRE_CTX(c);
  if (!conv_ratio || !iteration) { set_error("null argument"); return CLOTHGPU_ERR_ARG; }
  return run_pcg(c, conv_ratio[0], iteration[0], conv_ratio, iteration);
}
int clothgpu_cg(clothgpu_ctx* c, double* conv_ratio, int* iteration) {
  REQUIRE_CTX(c);
  if (!conv_ratio || !iteration) { set_error("null argument"); return CLOTHGPU_ERR_ARG; }
  return run_cg(c, conv_ratio[0], iteration[0], conv_ratio, iteration);
}

int clothgpu_step_internal_ilu(clothgpu_ctx* c, const clothgpu_params* p, double* conv_ratio, int* iteration) {
  REQUIRE_CTX(c);
  CG_CHECK(check_params(p));
  CG_CHECK(run_step_internal(c, *p, true, conv_ratio, iteration));
  return clothgpu_synchronize(c);
}
int clothgpu_step_internal_cg(clothgpu_ctx* c, const clothgpu_params* p, double* conv_ratio, int* iteration) {
  REQUIRE_CTX(c);
  CG_CHECK(check_params(p));
  CG_CHECK(run_step_internal(c, *p, false, conv_ratio, iteration));
  return clothgpu_synchronize(c);
}
int clothgpu_update_intermediate_velocity(clothgpu_ctx* c, const clothgpu_params* p) {
  REQUIRE_CTX(c);
  CG_CHECK(check_params(p));
  CG_CHECK(set_active(c, nullptr));
  CG_CHECK(launch_update_velocity(c, *p));
  return clothgpu_synchronize(c);
}

int clothgpu_bvh_refit_prx(clothgpu_ctx* c, int buf_xyz, double delta) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  double* x = need_vec(c, buf_xyz);
  if (!x) return CLOTHGPU_ERR_ARG;
  if (!(delta * 0.5 > 0)) { set_error("delta must be positive (CAABB3D::AddPoint ignores eps <= 0)"); return CLOTHGPU_ERR_ARG; }
  CG_CHECK(set_active(c, nullptr));
  CG_CHECK(launch_refit(c, x, nullptr, delta, 0));
  return clothgpu_synchronize(c);
}
int clothgpu_bvh_refit_ccd(clothgpu_ctx* c, int buf_xyz, int buf_uvw, double dt) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  double* x = need_vec(c, buf_xyz);
  double* v = need_vec(c, buf_uvw);
  if (!x || !v) return CLOTHGPU_ERR_ARG;
  CG_CHECK(set_active(c, nullptr));
  CG_CHECK(launch_refit(c, x, v, dt, 1));
  return clothgpu_synchronize(c);
}
int clothgpu_contacts_prx(clothgpu_ctx* c, int buf_xyz, double delta, int* counts) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  double* x = need_vec(c, buf_xyz);
  if (!x) return CLOTHGPU_ERR_ARG;
  CG_CHECK(set_active(c, nullptr));
  return run_contacts(c, x, nullptr, 0.0, delta, 0, counts);
}
int clothgpu_contacts_ccd(clothgpu_ctx* c, int buf_xyz, int buf_uvw, double dt, double delta, int* counts) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  double* x = need_vec(c, buf_xyz);
  double* v = need_vec(c, buf_uvw);
  if (!x || !v) return CLOTHGPU_ERR_ARG;
  CG_CHECK(set_active(c, nullptr));
  return run_contacts(c, x, v, dt, delta, 1, counts);
}
int clothgpu_leaf_pair_counts(clothgpu_ctx* c, int* counts) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  if (!counts) { set_error("null argument"); return CLOTHGPU_ERR_ARG; }
  CG_CHECK(fetch_counters(c));
  for (int s = 0; s < c->B; s++) counts[s] = c->h_cnt[s * NCNT + CNT_LEAF];
  return CLOTHGPU_OK;
}
int clothgpu_impulse_prx(clothgpu_ctx* c, int buf_xyz, int buf_uvw, double delta, double stiffness, double dt, double mass) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  double* x = need_vec(c, buf_xyz);
  double* v = need_vec(c, buf_uvw);
  if (!x || !v) return CLOTHGPU_ERR_ARG;
  CG_CHECK(fetch_counters(c));
  CG_CHECK(set_active(c, nullptr));
  CG_CHECK(launch_impulse(c, x, v, delta, stiffness, dt, mass, 0));
  return clothgpu_synchronize(c);
}
int clothgpu_impulse_ccd(clothgpu_ctx* c, int buf_xyz, int buf_uvw, double delta, double stiffness, double dt, double mass) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  double* x = need_vec(c, buf_xyz);
  double* v = need_vec(c, buf_uvw);
  if (!x || !v) return CLOTHGPU_ERR_ARG;
  CG_CHECK(fetch_counters(c));
  CG_CHECK(set_active(c, nullptr));
  CG_CHECK(launch_impulse(c, x, v, delta, stiffness, dt, mass, 1));
  return clothgpu_synchronize(c);
}
int clothgpu_riz_reset(clothgpu_ctx* c) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  CG_CHECK(launch_riz_reset(c));
  return clothgpu_synchronize(c);
}
int clothgpu_riz_make(clothgpu_ctx* c) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  CG_CHECK(fetch_counters(c));
  CG_CHECK(set_active(c, nullptr));
  CG_CHECK(launch_riz_make(c));
  return clothgpu_synchronize(c);
}
int clothgpu_riz_apply(clothgpu_ctx* c, int buf_xyz, int buf_uvw) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  double* x = need_vec(c, buf_xyz);
  double* v = need_vec(c, buf_uvw);
  if (!x || !v) return CLOTHGPU_ERR_ARG;
  CG_CHECK(fetch_counters(c));
  CG_CHECK(set_active(c, nullptr));
  CG_CHECK(launch_riz_apply(c, x, v));
  return clothgpu_synchronize(c);
}

int clothgpu_resolve_self_contact(clothgpu_ctx* c, const clothgpu_params* p, int* applied, int* stats) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  CG_CHECK(check_params(p));
  CG_CHECK(run_resolve(c, *p, applied, stats));
  return clothgpu_synchronize(c);
}

int clothgpu_step_self_contact(clothgpu_ctx* c, const clothgpu_params* p, int nsteps, int* stats) {
  REQUIRE_CTX(c);
  REQUIRE_BVH(c);
  CG_CHECK(check_params(p));
  std::vector<int> applied(c->B), iters(c->B);
  for (int step = 0; step < nsteps; step++) {
    // aXYZ1 = aXYZ (self_contact_sparse/main.cpp:224)
    CG_CUDA(cudaMemcpyAsync(c->xyz_prev, c->xyz, sizeof(double) * (size_t)c->B * c->ndof, cudaMemcpyDeviceToDevice, c->stream));
    // The proximity detection (:369-395) reads only aXYZ1: it runs on the second stream while the latency-bound ILU
    // sweeps of the implicit solve occupy 26 of the 148 SMs.  Measurement modes keep everything on one stream.
    const bool overlap = !c->ktiming && !c->profiling && p->contact_clearance * 0.5 > 0;
    if (overlap) {
      CG_CUDA(cudaEventRecord(c->ev_fork, c->stream));
      CG_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
      cudaStream_t main_stream = c->stream;
      c->stream = c->stream2;
      int rc = set_all_active(c, 1);
      if (rc == CLOTHGPU_OK) rc = launch_refit(c, c->xyz_prev, nullptr, p->contact_clearance, 0);
      if (rc == CLOTHGPU_OK) rc = launch_contacts_async(c, c->xyz_prev, nullptr, p->dt, p->contact_clearance, 0);
      c->stream = main_stream;
      CG_CHECK(rc);
      CG_CUDA(cudaEventRecord(c->ev_join, c->stream2));
    }
    {
      const int rc = run_step_internal(c, *p, true, nullptr, iters.data());
      // join the second stream on every path: its work must not outlive a failed step
      if (overlap) CG_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
      CG_CHECK(rc);
    }
    CG_CHECK(run_resolve(c, *p, applied.data(), stats, overlap));
    if (stats)
      for (int s = 0; s < c->B; s++) stats[s * CLOTHGPU_NSTATS + 0] = iters[s];
    bool any = false;
    for (int s = 0; s < c->B; s++) any = any || applied[s];
    if (any) {  // :246-261: re-integrate positions, then the half-step force correction
      CG_CHECK(set_active(c, applied.data()));
      phase_begin(c, 8);
      dim3 g(std::min(div_up(c->ndof, 256), 296), c->B);
      k_reintegrate<<<g, 256, 0, KS(c, "k_reintegrate")>>>(c->ndof, p->dt, c->xyz_prev, c->uvw, c->xyz, c->cnt);
      CG_LAUNCH_CHECK(c);
      CG_CHECK(launch_update_velocity(c, *p));
      phase_end(c, 8);
    }
  }
  return clothgpu_synchronize(c);
}

int clothgpu_step_internal(clothgpu_ctx* c, const clothgpu_params* p, int nsteps, int* iterations) {
  REQUIRE_CTX(c);
  CG_CHECK(check_params(p));
  for (int step = 0; step < nsteps; step++) CG_CHECK(run_step_internal(c, *p, true, nullptr, iterations));
  return clothgpu_synchronize(c);
}

}  // extern "C"
