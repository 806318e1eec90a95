// geo.cu -- the consumer right after predict(): great-circle error of the predicted class medians.
//
// Replaces the per-user Python loop of gcnmain.geo_eval (gcnmain.py:43-63): for every evaluated user the
// predicted class id is looked up in classLatMedian / classLonMedian and the haversine distance (km) to the
// user's true location is computed (third-party `haversine` package, unpinned in requirements.txt; its published
// formula: d = 2 R asin(sqrt(sin^2(dlat/2) + cos(lat1) cos(lat2) sin^2(dlon/2))), R = 6371.0088 km).
// float64 like the reference; one thread per user; HBM-bound (40 B in, 8 B out per user).
#include "common.cuh"

namespace {

__global__ void geo_distance_kernel(const long long* __restrict__ preds, int n, const double* __restrict__ class_lat,
                                    const double* __restrict__ class_lon, int n_classes,
                                    const double* __restrict__ lat_true, const double* __restrict__ lon_true,
                                    double radius_km, double* __restrict__ dist, int* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long c = preds[i];
  if (c < 0 || c >= n_classes) {
    *bad = 1;
    dist[i] = 0.0;
    return;
  }
  constexpr double kRad = 0.017453292519943295;  // pi / 180 (math.radians)
  // round-to-nearest intrinsics: no FMA contraction, so the expression rounds like the reference's Python floats
  // (and the distance of a point to itself is exactly 0)
  const double lat1 = __dmul_rn(lat_true[i], kRad), lon1 = __dmul_rn(lon_true[i], kRad);
  const double lat2 = __dmul_rn(class_lat[c], kRad), lon2 = __dmul_rn(class_lon[c], kRad);
  const double sl = sin(__dmul_rn(__dsub_rn(lat2, lat1), 0.5)), so = sin(__dmul_rn(__dsub_rn(lon2, lon1), 0.5));
  const double a = __dadd_rn(__dmul_rn(sl, sl), __dmul_rn(__dmul_rn(cos(lat1), cos(lat2)), __dmul_rn(so, so)));
  dist[i] = __dmul_rn(__dmul_rn(2.0, radius_km), asin(sqrt(a)));
}

}  // namespace

extern "C" int gcnb_geo_distance_f64(gcnb_ctx* ctx, const int64_t* preds, int32_t n, const double* class_lat,
                                     const double* class_lon, int32_t n_classes, const double* lat_true,
                                     const double* lon_true, double radius_km, double* dist, int32_t* bad_flag) {
  if (!ctx) return GCNB_E_INVALID;
  GCNB_REQUIRE(ctx, n >= 0 && n_classes > 0, "bad size");
  if (n == 0) return GCNB_OK;
  GCNB_REQUIRE(ctx, preds && class_lat && class_lon && lat_true && lon_true && dist && bad_flag, "null pointer");
  ProfScope scope(ctx, GCNB_TAG_LOSS);
  GCNB_CUDA(ctx, cudaMemsetAsync(bad_flag, 0, sizeof(int32_t), ctx->stream));
  geo_distance_kernel<<<cdiv(n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const long long*>(preds), n, class_lat,
                                                              class_lon, n_classes, lat_true, lon_true, radius_km, dist,
                                                              bad_flag);
  GCNB_LAUNCHED(ctx);
  return GCNB_OK;
}
