// calibration.cu -- the caller of the hot path: VisualCameraCalibration, NID_NELDER_MEAD branch
// (reference: src/vlcal/calib/visual_camera_calibration.cpp:35-139, include/vlcal/calib/visual_camera_calibration.hpp).
//
// Same decisions as the reference (outer loop, culling at the start pose of every inner solve, one cost object per
// bag, dfo::NelderMead<6> with the calibration parameters), but every Nelder-Mead iteration scores its candidate
// poses {xo, xr, xe, xc} in ONE batched kernel launch per bag, and bags run concurrently on their own streams.
#include <chrono>
#include <cstring>
#include <vector>

#include "host_math.hpp"
#include "mem_pool.hpp"
#include "nid_context.cuh"
#include "nid_kernels.cuh"

using namespace vlcal;

namespace {

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

vlcal_p2p* g_default_p2p = nullptr;  // vlcal_nid_p2p_set_default
int g_solver_mode = 0;                // vlcal_nid_set_solver_mode: 0 auto, 1 host loop, 2 device-resident loop

// the kernel sums over ranks only when a CONNECTED exchange spanning more than one rank is attached (fill_common_args)
inline bool ctx_exchange_is_fused(const vlcal_nid_ctx* c) {
  return c->p2p != nullptr && c->p2p->connected && c->p2p->world > 1;
}

struct PoseObjective {
  vlcal_nid_ctx* const* ctxs;
  int n_ctxs;
  const double* init_T;
  vlcal_pose_callback callback;
  vlcal_allreduce_fn allreduce;
  void* user;
  double best_cost = DBL_MAX;  // visual_camera_calibration.cpp:101
  int status = VLCAL_OK;
  std::vector<double> Ts, partial, vals;

  // f(x) for a batch: T = init_T * Expmap(x)  (:104), sum over bags of calculate(T)  (:105-110)
  void evaluate(const double* xs, int count, double* ys) {
    Ts.resize(static_cast<size_t>(count) * 16);
    partial.assign(count, 0.0);
    vals.resize(count);
    for (int i = 0; i < count; i++) {
      double E[16];
      host::se3_expmap_gtsam(xs + 6 * i, E);
      host::isometry_mul(init_T, E, &Ts[16 * static_cast<size_t>(i)]);
    }
    if (status == VLCAL_OK) {
      int launched = 0;
      for (; launched < n_ctxs; launched++) {
        const int rc = nid_evaluate_async(ctxs[launched], Ts.data(), count, false);
        if (rc != VLCAL_OK) {
          status = rc;
          break;
        }
      }
      for (int b = 0; b < launched; b++) {
        const int rc = nid_wait(ctxs[b], vals.data(), nullptr);
        if (rc != VLCAL_OK) {
          status = rc;
          continue;
        }
        for (int i = 0; i < count; i++) partial[i] += vals[i];  // sum_costs += costs[i]->calculate(T)
      }
    }
    // with a peer exchange attached (one bag per rank) the kernel already returned the sum over ranks
    const bool fused = n_ctxs == 1 && ctx_exchange_is_fused(ctxs[0]);
    if (allreduce && !fused) allreduce(partial.data(), count, user);
    for (int i = 0; i < count; i++) ys[i] = status == VLCAL_OK ? partial[i] : NAN;
  }

  // the objective's side effects, in the reference's evaluation order (:112-116)
  void observe(const double* x, double y) {
    if (y < best_cost) {
      best_cost = y;
      if (callback) {
        double E[16], T[16];
        host::se3_expmap_gtsam(x, E);
        host::isometry_mul(init_T, E, T);
        callback(T, y, user);
      }
    }
  }
};

// Device-resident inner solve: the Nelder-Mead state machine advances inside the finalizing block of each launch
// (nid_kernels.cuh: nm_device_advance), the host only enqueues launches in chunks and replays the best-cost callback
// from the evaluation trace afterwards (same order as the reference, delivered up to one chunk late).
constexpr int DEVICE_LOOP_CHUNK = 24;

struct PoolPtr {
  void* p = nullptr;
  int device = 0;
  bool pinned = false;
  ~PoolPtr() {
    if (pinned) MemPool::instance().pinned_free(p);
    else MemPool::instance().device_free(device, p);
  }
};

int run_inner_solve_device(
  vlcal_nid_ctx* ctx, const vlcal_calib_params* params, const double init_T[16], vlcal_pose_callback callback, void* user, double T_out[16], vlcal_nm_result* nm_result) {
  VL_CUDA(cudaSetDevice(ctx->device));
  NmParams nm;  // visual_camera_calibration.cpp:122-125
  nm.init_step = params->nelder_mead_init_step;
  nm.convergence_var_thresh = params->nelder_mead_convergence_criteria;
  nm.max_iterations = params->max_inner_iterations;

  PoolPtr h_state, d_state, d_trace, h_trace;
  h_state.pinned = true, h_trace.pinned = true;
  d_state.device = d_trace.device = ctx->device;
  const int trace_cap = 16 + std::max(0, params->max_inner_iterations) * (NM_MAX_N + 1);
  const size_t trace_bytes = sizeof(double) * (NM_MAX_N + 1) * static_cast<size_t>(trace_cap);
  VL_CUDA(MemPool::instance().pinned_alloc(sizeof(NmDevice), &h_state.p));
  VL_CUDA(MemPool::instance().device_alloc(ctx->device, sizeof(NmDevice), &d_state.p));
  VL_CUDA(MemPool::instance().device_alloc(ctx->device, trace_bytes, &d_trace.p));
  VL_CUDA(MemPool::instance().pinned_alloc(trace_bytes, &h_trace.p));

  NmDevice* h = static_cast<NmDevice*>(h_state.p);
  std::memset(static_cast<void*>(h), 0, sizeof(NmDevice));
  const double x0[6] = {0, 0, 0, 0, 0, 0};
  h->nm.begin(6, nm, x0);  // :126 optimize(f, Zero)
  std::memcpy(h->init_T, init_T, sizeof(h->init_T));
  for (int k = 0; k < h->nm.n_cand; k++) nm_pose_of_candidate(h->nm.cand[k], h->init_T, h, k);
  h->n_poses = h->nm.n_cand;
  h->trace_cap = trace_cap;
  h->trace_count = 0;
  h->trace = static_cast<double*>(d_trace.p);
  VL_CUDA(cudaMemcpyAsync(d_state.p, h, sizeof(NmDevice), cudaMemcpyHostToDevice, ctx->stream));

  double best_cost = DBL_MAX;  // :101
  int trace_seen = 0;
  unsigned long long steps_seen = 0;
  int computed_seen = 0;
  const double* trace = static_cast<const double*>(h_trace.p);
  for (;;) {
    int rc = nid_enqueue_device_steps(ctx, static_cast<NmDevice*>(d_state.p), DEVICE_LOOP_CHUNK);
    if (rc != VLCAL_OK) return rc;
    VL_CUDA(cudaMemcpyAsync(h, d_state.p, sizeof(NmDevice), cudaMemcpyDeviceToHost, ctx->stream));
    VL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->p2p && ctx->p2p->h_error && *ctx->p2p->h_error) {
      set_last_error("peer exchange timed out: a rank died or the ranks are not evaluating in lockstep");
      return VLCAL_ERR_CUDA;
    }
    rc = nid_account_device_steps(ctx, DEVICE_LOOP_CHUNK, static_cast<int>(h->steps_done - steps_seen), h->nm.num_evaluations_computed - computed_seen);
    if (rc != VLCAL_OK) return rc;
    steps_seen = h->steps_done;
    computed_seen = h->nm.num_evaluations_computed;
    const int count = std::min(h->trace_count, trace_cap);
    if (count > trace_seen) {  // replay the objective's side effects in the reference's order (:112-116)
      const size_t stride = NM_MAX_N + 1;
      VL_CUDA(cudaMemcpy(const_cast<double*>(trace) + stride * trace_seen, static_cast<double*>(d_trace.p) + stride * trace_seen, sizeof(double) * stride * (count - trace_seen), cudaMemcpyDeviceToHost));
      for (int k = trace_seen; k < count; k++) {
        const double* e = trace + stride * k;
        const double y = e[NM_MAX_N];
        if (y < best_cost) {
          best_cost = y;
          if (callback) {
            double E[16], T[16];
            host::se3_expmap_gtsam(e, E);
            host::isometry_mul(init_T, E, T);
            callback(T, y, user);
          }
        }
      }
      trace_seen = count;
    }
    if (h->nm.phase == 3) break;
    if (h->steps_done == steps_seen && h->n_poses == 0) break;  // defensive: nothing left to do
  }
  double E[16];
  host::se3_expmap_gtsam(h->nm.result_x, E);
  host::isometry_mul(init_T, E, T_out);  // :129
  if (nm_result) {
    std::memset(nm_result, 0, sizeof(*nm_result));
    nm_result->converged = h->nm.converged;
    nm_result->num_iterations = h->nm.num_iterations;
    for (int d = 0; d < 6; d++) nm_result->x[d] = h->nm.result_x[d];
    nm_result->y = h->nm.result_y;
    nm_result->num_evaluations = h->nm.num_evaluations;
    nm_result->num_batches = h->nm.num_batches;
    nm_result->num_evaluations_computed = h->nm.num_evaluations_computed;
  }
  return VLCAL_OK;
}

int run_inner_solve(
  vlcal_nid_ctx* const* ctxs, int n_ctxs, const vlcal_calib_params* params, const double init_T[16], vlcal_pose_callback callback, vlcal_allreduce_fn allreduce, void* user,
  double T_out[16], vlcal_nm_result* nm_result) {
  // Persistent cooperative kernel (nid_persistent.cuh; default): the whole solve is ONE launch -- candidates scored,
  // summed over bags and ranks, and the Nelder-Mead machine stepped inside the kernel.  Needs what pk_supported checks
  // (float4 clouds, lean classifier, bins <= 32, one camera / image size / device) and no host-side all-reduce callback
  // (a connected peer exchange is summed in-kernel instead).
  bool same_px = true;
  for (int i = 1; i < n_ctxs; i++) same_px = same_px && ctxs[i]->p2p == ctxs[0]->p2p;
  const bool px_fused = n_ctxs >= 1 && ctx_exchange_is_fused(ctxs[0]);
  const bool pk_ok = n_ctxs >= 1 && params->max_inner_iterations >= 0 && same_px && (allreduce == nullptr || px_fused) && pk_supported(ctxs, n_ctxs);
  if (g_solver_mode == 3 && !pk_ok) {
    set_last_error("persistent solve requested but these contexts need the host loop (double-layout cloud / bins > 32 / camera without the lean classifier / host all-reduce / mixed image sizes)");
    return VLCAL_ERR_UNSUPPORTED;
  }
  if (pk_ok && (g_solver_mode == 0 || g_solver_mode == 3)) {
    return pk_solve(ctxs, n_ctxs, params, init_T, callback, user, T_out, nm_result);
  }
  // one bag on this GPU, scores either local or summed in-kernel over the peer exchange: the round-1 device-resident loop
  // (one launch per batch, machine stepped by the finalizing block).  (Several local bags, or a host-side all-reduce
  // callback, need the host between batches.)
  const bool device_ok = n_ctxs == 1 && ctxs[0]->mode == VLCAL_NID_MODE_HISTOGRAM && ctxs[0]->max_poses >= NID_MAX_POSES && (allreduce == nullptr || ctx_exchange_is_fused(ctxs[0])) &&
                         params->max_inner_iterations >= 0;
  if (g_solver_mode == 2 && !device_ok) {
    set_last_error("device-resident solver loop requested but this solve needs the host between batches (several local bags / host all-reduce / bins too large)");
    return VLCAL_ERR_UNSUPPORTED;
  }
  if (device_ok && g_solver_mode == 2) {
    return run_inner_solve_device(ctxs[0], params, init_T, callback, user, T_out, nm_result);
  }
  if (n_ctxs > 1 && px_fused) {
    // the round-1 kernels exchange one bag per rank; with several local bags only the persistent kernel sums in-kernel
    set_last_error("several local bags with an attached peer exchange need the persistent solve (solver mode 0 / 3)");
    return VLCAL_ERR_UNSUPPORTED;
  }
  PoseObjective obj;
  obj.ctxs = ctxs, obj.n_ctxs = n_ctxs, obj.init_T = init_T, obj.callback = callback, obj.allreduce = allreduce, obj.user = user;

  host::NelderMeadParams nm;  // :122-125
  nm.init_step = params->nelder_mead_init_step;
  nm.convergence_var_thresh = params->nelder_mead_convergence_criteria;
  nm.max_iterations = params->max_inner_iterations;
  const double x0[6] = {0, 0, 0, 0, 0, 0};
  auto f = [&](const double* xs, int count, double* ys) { obj.evaluate(xs, count, ys); };
  auto ob = [&](const double* x, double y) { obj.observe(x, y); };
  const host::NelderMeadResult r = host::nelder_mead(6, f, ob, x0, nm, /*speculate=*/true);  // :126-127
  if (obj.status != VLCAL_OK) return obj.status;

  double E[16];
  host::se3_expmap_gtsam(r.x.data(), E);
  host::isometry_mul(init_T, E, T_out);  // :129
  if (nm_result) {
    std::memset(nm_result, 0, sizeof(*nm_result));
    nm_result->converged = r.converged ? 1 : 0;
    nm_result->num_iterations = r.num_iterations;
    for (int d = 0; d < 6; d++) nm_result->x[d] = r.x[d];
    nm_result->y = r.y;
    nm_result->num_evaluations = r.num_evaluations;
    nm_result->num_batches = r.num_batches;
    nm_result->num_evaluations_computed = r.num_evaluations_computed;
  }
  return VLCAL_OK;
}

// a bag resident in HBM for the whole calibration (uploaded once, culled per outer iteration)
struct ResidentBag {
  std::shared_ptr<DeviceCloud> cloud;
  std::shared_ptr<DeviceImage> image;
  double max_fov = 0.0;  // estimate_camera_fov(proj, this bag's image size)  (cost_calculator_nid.cpp:17)
};

int upload_bags(int device, const CameraParams& cam, const vlcal_bag* bags, int n_bags, std::vector<ResidentBag>* out) {
  out->resize(n_bags);
  for (int b = 0; b < n_bags; b++) {
    const vlcal_bag& bag = bags[b];
    if (!bag.image || bag.width <= 0 || bag.height <= 0 || bag.row_stride_bytes < bag.width || bag.n_points < 0 || (bag.n_points > 0 && (!bag.points_xyzw || !bag.intensities))) {
      set_last_error("invalid bag " + std::to_string(b));
      return VLCAL_ERR_INVALID_ARGUMENT;
    }
    int rc = upload_cloud(device, bag.points_xyzw, bag.intensities, bag.n_points, nullptr, &(*out)[b].cloud);
    if (rc != VLCAL_OK) return rc;
    rc = upload_image(device, bag.image, bag.width, bag.height, bag.row_stride_bytes, nullptr, &(*out)[b].image);
    if (rc != VLCAL_OK) return rc;
    // identical for bags that share an image size; cheap (three 2-D Nelder-Mead solves on the host)
    (*out)[b].max_fov = (b > 0 && bag.width == bags[0].width && bag.height == bags[0].height) ? (*out)[0].max_fov : estimate_camera_fov_host(cam, bag.width, bag.height);
  }
  return VLCAL_OK;
}

struct CtxList {
  std::vector<vlcal_nid_ctx*> v;
  ~CtxList() {
    for (auto* c : v) delete c;
  }
};

// estimate_pose_nelder_mead (:70-139) on resident bags
int inner_solve_resident(
  int device, const CameraParams& cam, const std::vector<ResidentBag>& bags, const vlcal_calib_params* params, const double init_T[16], vlcal_pose_callback callback,
  vlcal_allreduce_fn allreduce, void* user, int profiling, double T_out[16], vlcal_nm_result* nm_result, vlcal_calib_stats* stats, int outer_index) {
  // :71-73 one ViewCulling object, built on dataset.front()'s image size
  const double cull_fov = bags.empty() ? 0.0 : bags[0].max_fov;
  const double t_cull0 = now_ms();
  CtxList ctxs;
  for (size_t b = 0; b < bags.size(); b++) {
    std::shared_ptr<DeviceCloud> culled;
    int64_t kept = 0;
    int rc = view_cull_device(cam, bags[0].image->width, bags[0].image->height, cull_fov, !params->disable_z_buffer_culling, *bags[b].cloud, init_T, nullptr, &culled, nullptr, &kept);  // :78
    if (rc != VLCAL_OK) return rc;
    if (stats && b == 0 && outer_index < 16) stats->culled_points[outer_index] = kept;
    vlcal_nid_ctx* ctx = nullptr;
    rc = nid_ctx_create(device, VLCAL_NID_MODE_HISTOGRAM, cam, bags[b].image, culled, params->nid_bins, bags[b].max_fov, &ctx);  // :82-84
    if (rc != VLCAL_OK) return rc;
    ctx->profiling = profiling != 0;
    ctxs.v.push_back(ctx);
  }
  if (g_default_p2p && g_default_p2p->device == device) {
    // one bag per rank: every path can sum in-kernel; several local bags: only the persistent solve does
    const bool pk_path = (g_solver_mode == 0 || g_solver_mode == 3) && pk_supported(ctxs.v.data(), static_cast<int>(ctxs.v.size()));
    if (bags.size() == 1 || pk_path)
      for (auto* c : ctxs.v) c->p2p = g_default_p2p;
  }
  const double t_solve0 = now_ms();
  if (stats) stats->cull_ms += t_solve0 - t_cull0;
  vlcal_nm_result local;
  const int rc = run_inner_solve(ctxs.v.data(), static_cast<int>(ctxs.v.size()), params, init_T, callback, allreduce, user, T_out, &local);
  if (rc != VLCAL_OK) return rc;
  if (nm_result) *nm_result = local;
  if (stats) {
    stats->solve_ms += now_ms() - t_solve0;
    stats->total_evaluations += local.num_evaluations;
    stats->total_evaluations_computed += local.num_evaluations_computed;
    stats->total_batches += local.num_batches;
    if (outer_index < 16) {
      stats->inner_iterations[outer_index] = local.num_iterations;
      stats->inner_final_cost[outer_index] = local.y;
    }
    for (auto* c : ctxs.v) {
      int64_t launches = 0, poses = 0;
      double ms = 0.0;
      if (vlcal_nid_get_profile(c, &launches, &ms, &poses) == VLCAL_OK) {
        stats->kernel_launches += launches;
        stats->kernel_ms_total += ms;
      }
    }
  }
  return VLCAL_OK;
}

int check_common(const vlcal_calib_params* params, const double* init_T, double* T_out, int n_bags) {
  if (!params || !init_T || !T_out || n_bags <= 0) {
    set_last_error("invalid arguments");
    return VLCAL_ERR_INVALID_ARGUMENT;
  }
  if (vlcal_nid_device_count() == 0) {
    set_last_error("no CUDA device available: this library has no CPU fallback");
    return VLCAL_ERR_NO_DEVICE;
  }
  return VLCAL_OK;
}

}  // namespace

extern "C" {

int vlcal_nid_set_solver_mode(int mode) {
  if (mode < 0 || mode > 3) {
    set_last_error("solver mode must be 0 (auto), 1 (host loop), 2 (device-resident loop) or 3 (persistent kernel)");
    return VLCAL_ERR_INVALID_ARGUMENT;
  }
  g_solver_mode = mode;
  return VLCAL_OK;
}

int vlcal_nid_p2p_set_default(vlcal_p2p* px) {
  if (px && !px->connected) {
    set_last_error("peer exchange is not connected");
    return VLCAL_ERR_INVALID_ARGUMENT;
  }
  g_default_p2p = px;
  return VLCAL_OK;
}

void vlcal_nm_default_params(vlcal_nm_params* p) {  // nelder_mead.hpp:12
  const host::NelderMeadParams d;
  p->init_step = d.init_step;
  p->alpha = d.alpha;
  p->gamma = d.gamma;
  p->rho = d.rho;
  p->sigma = d.sigma;
  p->max_iterations = d.max_iterations;
  p->convergence_var_thresh = d.convergence_var_thresh;
}

int vlcal_nelder_mead_batched(int n, vlcal_nm_batch_fn f, vlcal_nm_observe_fn observe, void* user, const double* x0, const vlcal_nm_params* params, vlcal_nm_result* result) {
  if (n < 1 || n > NM_MAX_N || !f || !x0 || !params || !result) {
    set_last_error("invalid arguments (1 <= n <= 8)");
    return VLCAL_ERR_INVALID_ARGUMENT;
  }
  host::NelderMeadParams p;
  p.init_step = params->init_step;
  p.alpha = params->alpha;
  p.gamma = params->gamma;
  p.rho = params->rho;
  p.sigma = params->sigma;
  p.max_iterations = params->max_iterations;
  p.convergence_var_thresh = params->convergence_var_thresh;
  auto bf = [&](const double* xs, int count, double* ys) { f(xs, count, n, ys, user); };
  auto ob = [&](const double* x, double y) {
    if (observe) observe(x, n, y, user);
  };
  const host::NelderMeadResult r = host::nelder_mead(n, bf, ob, x0, p, /*speculate=*/true);
  std::memset(result, 0, sizeof(*result));
  result->converged = r.converged ? 1 : 0;
  result->num_iterations = r.num_iterations;
  for (int d = 0; d < n; d++) result->x[d] = r.x[d];
  result->y = r.y;
  result->num_evaluations = r.num_evaluations;
  result->num_batches = r.num_batches;
  result->num_evaluations_computed = r.num_evaluations_computed;
  return VLCAL_OK;
}

void vlcal_calib_default_params(vlcal_calib_params* p) {  // visual_camera_calibration.hpp:12-26
  p->max_outer_iterations = 10;
  p->max_inner_iterations = 256;
  p->delta_trans_thresh = 0.1;
  p->delta_rot_thresh = 0.5 * M_PI / 180.0;
  p->disable_z_buffer_culling = 0;
  p->nid_bins = 16;
  p->nelder_mead_init_step = 1e-3;
  p->nelder_mead_convergence_criteria = 1e-8;
}

int vlcal_estimate_pose_nelder_mead_ctx(
  vlcal_nid_ctx* const* ctxs,
  int n_ctxs,
  const vlcal_calib_params* params,
  const double init_T_camera_lidar[16],
  vlcal_pose_callback callback,
  vlcal_allreduce_fn allreduce,
  void* user,
  double T_out[16],
  vlcal_nm_result* nm_result) {
  if (!ctxs || n_ctxs < 0 || !params || !init_T_camera_lidar || !T_out) {
    set_last_error("invalid arguments");
    return VLCAL_ERR_INVALID_ARGUMENT;
  }
  for (int i = 0; i < n_ctxs; i++) {
    if (!ctxs[i]) {
      set_last_error("NULL context");
      return VLCAL_ERR_INVALID_ARGUMENT;
    }
  }
  return run_inner_solve(ctxs, n_ctxs, params, init_T_camera_lidar, callback, allreduce, user, T_out, nm_result);
}

int vlcal_estimate_pose_nelder_mead(
  int device,
  int camera_model,
  const double* intrinsics,
  int n_intrinsics,
  const double* distortion,
  int n_distortion,
  const vlcal_bag* bags,
  int n_bags,
  const vlcal_calib_params* params,
  const double init_T_camera_lidar[16],
  vlcal_pose_callback callback,
  vlcal_allreduce_fn allreduce,
  void* user,
  int profiling,
  double T_out[16],
  vlcal_nm_result* nm_result,
  vlcal_calib_stats* stats) {
  int rc = check_common(params, init_T_camera_lidar, T_out, n_bags);
  if (rc != VLCAL_OK) return rc;
  CameraParams cam;
  rc = make_camera(camera_model, intrinsics, n_intrinsics, distortion, n_distortion, &cam);
  if (rc != VLCAL_OK) return rc;
  if (device < 0) VL_CUDA(cudaGetDevice(&device));
  VL_CUDA(cudaSetDevice(device));
  std::vector<ResidentBag> resident;
  const double t_up0 = now_ms();
  rc = upload_bags(device, cam, bags, n_bags, &resident);
  if (rc != VLCAL_OK) return rc;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->upload_ms = now_ms() - t_up0;
  }
  rc = inner_solve_resident(device, cam, resident, params, init_T_camera_lidar, callback, allreduce, user, profiling, T_out, nm_result, stats, 0);
  if (rc == VLCAL_OK && stats) stats->outer_iterations = 1;
  return rc;
}

int vlcal_calibrate_nelder_mead(
  int device,
  int camera_model,
  const double* intrinsics,
  int n_intrinsics,
  const double* distortion,
  int n_distortion,
  const vlcal_bag* bags,
  int n_bags,
  const vlcal_calib_params* params,
  const double init_T_camera_lidar[16],
  vlcal_pose_callback callback,
  vlcal_allreduce_fn allreduce,
  void* user,
  int profiling,
  double T_out[16],
  vlcal_calib_stats* stats) {
  int rc = check_common(params, init_T_camera_lidar, T_out, n_bags);
  if (rc != VLCAL_OK) return rc;
  CameraParams cam;
  rc = make_camera(camera_model, intrinsics, n_intrinsics, distortion, n_distortion, &cam);
  if (rc != VLCAL_OK) return rc;
  if (device < 0) VL_CUDA(cudaGetDevice(&device));
  VL_CUDA(cudaSetDevice(device));
  std::vector<ResidentBag> resident;
  const double t_up0 = now_ms();
  rc = upload_bags(device, cam, bags, n_bags, &resident);
  if (rc != VLCAL_OK) return rc;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->upload_ms = now_ms() - t_up0;
  }

  double T[16];
  std::memcpy(T, init_T_camera_lidar, sizeof(T));
  for (int i = 0; i < params->max_outer_iterations; i++) {  // visual_camera_calibration.cpp:39
    double new_T[16];
    rc = inner_solve_resident(device, cam, resident, params, T, callback, allreduce, user, profiling, new_T, nullptr, stats, i);  // :46
    if (rc != VLCAL_OK) return rc;
    double inv[16], delta[16];
    host::isometry_inverse(new_T, inv);
    host::isometry_mul(inv, T, delta);  // :50
    std::memcpy(T, new_T, sizeof(T));   // :51
    const double tx = host::m4(delta, 0, 3), ty = host::m4(delta, 1, 3), tz = host::m4(delta, 2, 3);
    const double delta_t = std::sqrt((tx * tx + ty * ty) + tz * tz);  // :53
    const double delta_r = host::rotation_angle(delta);               // :54
    const bool converged = delta_t < params->delta_trans_thresh && delta_r < params->delta_rot_thresh;  // :55
    if (stats) stats->outer_iterations = i + 1;
    if (converged) break;  // :62-64
  }
  std::memcpy(T_out, T, sizeof(T));
  return VLCAL_OK;
}

}  // extern "C"
