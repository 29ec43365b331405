// XLA-FFI adapter: the launchers of include/b200ppo.h as typed XLA custom calls, so that the reference's
// own jitted JAX program (nnx_ppo/algorithms/ppo.py) can call the B200 kernels through jax.ffi.ffi_call
// (BASELINE.json north_star: "a thin XLA-FFI (jax.ffi custom-call) C-ABI layer").
//
// Built by nnx_ppo_b200/xla_ffi.py::build() when `jax.ffi.include_dir()` resolves (JAX is not installed
// in the image this repository is developed in, so the file is compiled only where JAX exists):
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include nnx_ppo_b200/csrc/xla/b200ppo_xla.cc \
//       -L nnx_ppo_b200/lib -lb200ppo -Wl,-rpath,'$ORIGIN' -o nnx_ppo_b200/lib/libb200ppo_xla.so
//
// Conventions: every handler takes the XLA stream, forwards raw device pointers, never allocates or
// synchronises (the launchers do not either), and maps a non-zero return code to ffi::Error.  Buffers
// that the C entry point updates in place are declared as operands AND results and must be aliased by
// the caller (`input_output_aliases` of jax.ffi.ffi_call); the handler checks the aliasing.  POD structs
// (b200ppo_plan, b200ppo_hparams) travel as byte-array attributes with the exact C layout
// (nnx_ppo_b200/_lib.py holds the ctypes mirror the Python side serialises).
#include <cstdint>
#include <cstring>

#include <cuda_runtime_api.h>

#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

#include "b200ppo.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error status(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + b200ppo_error_string(rc));
}

template <class T>
bool pod_from(ffi::Span<const uint8_t> bytes, T* out) {
  if (bytes.size() != sizeof(T)) return false;
  std::memcpy(out, bytes.begin(), sizeof(T));
  return true;
}

// ---- gae (ppo.py:351-394): [T, B] rewards / values, [B] last value, [T, B] bool masks -> [T, B] ----
ffi::Error GaeImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> rewards, ffi::Buffer<ffi::F32> values,
                   ffi::Buffer<ffi::F32> last_value, ffi::Buffer<ffi::PRED> done, ffi::Buffer<ffi::PRED> truncation,
                   float lambda_, float gamma, ffi::ResultBuffer<ffi::F32> advantages) {
  const auto d = rewards.dimensions();
  if (d.size() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200ppo_gae: rewards must be [T, B]");
  if (values.element_count() != rewards.element_count() || done.element_count() != rewards.element_count() ||
      truncation.element_count() != rewards.element_count() || last_value.element_count() != static_cast<size_t>(d[1]) ||
      advantages->element_count() != rewards.element_count())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200ppo_gae: shape mismatch");
  return status(b200ppo_gae(stream, rewards.typed_data(), values.typed_data(), last_value.typed_data(),
                            reinterpret_cast<const uint8_t*>(done.typed_data()),
                            reinterpret_cast<const uint8_t*>(truncation.typed_data()), static_cast<int32_t>(d[0]),
                            static_cast<int32_t>(d[1]), lambda_, gamma, advantages->typed_data()),
                "b200ppo_gae");
}

// ---- minibatch permutation indices (ppo.py:287-294): key data uint32[2] -> int32 [n_epochs, n] ----
ffi::Error PermutationImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> new_key, ffi::ResultBuffer<ffi::S32> indices,
                           ffi::ResultBuffer<ffi::U8> scratch) {
  const auto d = indices->dimensions();
  if (d.size() != 2 || new_key.element_count() != 2)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200ppo_permutation: key data uint32[2] -> indices [n_epochs, n]");
  const int32_t n_epochs = static_cast<int32_t>(d[0]), n = static_cast<int32_t>(d[1]);
  if (static_cast<int64_t>(scratch->element_count()) < b200ppo_permutation_scratch_bytes(n, n_epochs))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200ppo_permutation: scratch smaller than b200ppo_permutation_scratch_bytes");
  return status(b200ppo_permutation(stream, new_key.typed_data(), n, n_epochs, indices->typed_data(), scratch->typed_data()),
                "b200ppo_permutation");
}

// ---- one minibatch update (ppo.py:296-317 update_step = nnx.grad(ppo_loss) + optimizer.update) ----
// Operands (all device buffers): the time-major rollout record, the minibatch indices, Normalizer
// statistics, then the in-place state (params, adam mu, adam nu, workspace) which is aliased to the
// results of the same names; `counters` = uint32[4] (sampler stream key, sampler count, adam count).
ffi::Error UpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> obs, ffi::Buffer<ffi::F32> raw_action,
                      ffi::Buffer<ffi::F32> loglik_old, ffi::Buffer<ffi::F32> reward, ffi::Buffer<ffi::PRED> done,
                      ffi::Buffer<ffi::PRED> truncated, ffi::Buffer<ffi::F32> next_obs_last, ffi::Buffer<ffi::S32> inds,
                      ffi::Buffer<ffi::F32> norm_mean, ffi::Buffer<ffi::F32> norm_std, ffi::Buffer<ffi::U32> counters,
                      ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> adam_mu, ffi::Buffer<ffi::F32> adam_nu,
                      ffi::Buffer<ffi::F32> workspace, ffi::Span<const uint8_t> plan_bytes,
                      ffi::Span<const uint8_t> hparams_bytes, int32_t rng_count_offset, int32_t update_index,
                      int32_t stages, ffi::ResultBuffer<ffi::F32> params_out, ffi::ResultBuffer<ffi::F32> adam_mu_out,
                      ffi::ResultBuffer<ffi::F32> adam_nu_out, ffi::ResultBuffer<ffi::F32> workspace_out,
                      ffi::ResultBuffer<ffi::F32> metrics) {
  b200ppo_plan plan;
  b200ppo_hparams hp;
  if (!pod_from(plan_bytes, &plan) || !pod_from(hparams_bytes, &hp))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200ppo_update: plan / hparams attribute has the wrong size");
  if (params_out->untyped_data() != params.untyped_data() || adam_mu_out->untyped_data() != adam_mu.untyped_data() ||
      adam_nu_out->untyped_data() != adam_nu.untyped_data() || workspace_out->untyped_data() != workspace.untyped_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument,
                      "b200ppo_update: params / adam_mu / adam_nu / workspace must be aliased to their results "
                      "(input_output_aliases)");
  const auto d = obs.dimensions();
  if (d.size() != 3) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200ppo_update: obs must be [T, B, O]");
  if (metrics->element_count() < B200PPO_METRICS_STRIDE)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200ppo_update: metrics needs B200PPO_METRICS_STRIDE floats");
  const int32_t T = static_cast<int32_t>(d[0]), B = static_cast<int32_t>(d[1]);
  const int32_t mb = static_cast<int32_t>(inds.element_count());
  if (static_cast<int64_t>(workspace.element_count()) * 4 < b200ppo_update_workspace_bytes(&plan, T, mb))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "b200ppo_update: workspace smaller than b200ppo_update_workspace_bytes");
  b200ppo_update_bufs b;
  std::memset(&b, 0, sizeof(b));
  b.obs = obs.typed_data(); b.raw_action = raw_action.typed_data(); b.loglik_old = loglik_old.typed_data();
  b.reward = reward.typed_data();
  b.done = reinterpret_cast<const uint8_t*>(done.typed_data());
  b.truncated = reinterpret_cast<const uint8_t*>(truncated.typed_data());
  b.next_obs_last = next_obs_last.typed_data(); b.inds = inds.typed_data();
  b.norm_mean = norm_mean.typed_data(); b.norm_std = norm_std.typed_data();
  b.params = params_out->typed_data(); b.adam_mu = adam_mu_out->typed_data(); b.adam_nu = adam_nu_out->typed_data();
  b.rng_state = counters.typed_data();
  b.metrics_out = metrics->typed_data();
  b.ws = workspace_out->typed_data();
  return status(b200ppo_update(stream, &plan, &hp, &b, T, B, mb, static_cast<uint32_t>(rng_count_offset), update_index, stages),
                "b200ppo_update");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200ppoGae, GaeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()    // rewards            [T, B]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // values_excl_last   [T, B]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // last_value         [B]
                                  .Arg<ffi::Buffer<ffi::PRED>>()   // done               [T, B]
                                  .Arg<ffi::Buffer<ffi::PRED>>()   // truncation         [T, B]
                                  .Attr<float>("lambda_")
                                  .Attr<float>("gamma")
                                  .Ret<ffi::Buffer<ffi::F32>>());  // advantages         [T, B]

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200ppoPermutation, PermutationImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()    // jax.random.key_data(new_key)
                                  .Ret<ffi::Buffer<ffi::S32>>()    // indices [n_epochs, n]
                                  .Ret<ffi::Buffer<ffi::U8>>());   // scratch

XLA_FFI_DEFINE_HANDLER_SYMBOL(B200ppoUpdate, UpdateImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()    // obs            [T, B, O]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // raw_action     [T, B, A]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // loglik_old     [T, B]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // reward         [T, B]
                                  .Arg<ffi::Buffer<ffi::PRED>>()   // done           [T, B]
                                  .Arg<ffi::Buffer<ffi::PRED>>()   // truncated      [T, B]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // next_obs_last  [B, O]
                                  .Arg<ffi::Buffer<ffi::S32>>()    // inds           [mb]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // norm_mean      [O]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // norm_std       [O]
                                  .Arg<ffi::Buffer<ffi::U32>>()    // counters       [4]
                                  .Arg<ffi::Buffer<ffi::F32>>()    // params         [P]   (aliased)
                                  .Arg<ffi::Buffer<ffi::F32>>()    // adam_mu        [P]   (aliased)
                                  .Arg<ffi::Buffer<ffi::F32>>()    // adam_nu        [P]   (aliased)
                                  .Arg<ffi::Buffer<ffi::F32>>()    // workspace            (aliased)
                                  .Attr<ffi::Span<const uint8_t>>("plan")
                                  .Attr<ffi::Span<const uint8_t>>("hparams")
                                  .Attr<int32_t>("rng_count_offset")
                                  .Attr<int32_t>("update_index")
                                  .Attr<int32_t>("stages")
                                  .Ret<ffi::Buffer<ffi::F32>>()    // params
                                  .Ret<ffi::Buffer<ffi::F32>>()    // adam_mu
                                  .Ret<ffi::Buffer<ffi::F32>>()    // adam_nu
                                  .Ret<ffi::Buffer<ffi::F32>>()    // workspace
                                  .Ret<ffi::Buffer<ffi::F32>>());  // metrics [B200PPO_METRICS_STRIDE]
