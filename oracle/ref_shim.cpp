// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" veneer over the reference's *unmodified* C++ arithmetic core,
// compiled in place from /root/reference by oracle/Makefile into oracle/_ref/.
// The reference's own CPython wrappers (cpp_neighbors/wrapper.cpp,
// cpp_subsampling/wrapper.cpp) do not build against NumPy 2.x, so this shim takes
// their place: it does what wrapper.cpp:188-224 / wrapper.cpp:232-300 do (copy
// the flat arrays into std::vector<PointXYZ>, call the core, copy the result out)
// with plain pointers instead of PyArrayObjects.
//
// Entry points wrapped:
//   batch_nanoflann_neighbors  cpp_neighbors/neighbors/neighbors.cpp:211-332 (the one the reference wires in, wrapper.cpp:198)
//   batch_ordered_neighbors    cpp_neighbors/neighbors/neighbors.cpp:125-208 (brute force)
//   batch_grid_subsampling     cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211
#include "cpp_neighbors/neighbors/neighbors.h"
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"

#include <cstring>
#include <vector>

namespace {
std::vector<PointXYZ> to_points(const float* xyz, int n) {
  std::vector<PointXYZ> v(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) v[i] = PointXYZ(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  return v;
}
std::vector<int> g_nb;        // result of the last neighbour query
std::vector<PointXYZ> g_sub;  // result of the last subsampling
std::vector<int> g_sub_len;
}  // namespace

extern "C" {

// Returns max_count (row width); fetch the [nq, max_count] matrix with ref_neighbors_fetch.
int ref_batch_neighbors(const float* q, int nq, const float* s, int ns, const int* qb, const int* sb,
                        int nb, float radius, int brute) {
  std::vector<PointXYZ> qv = to_points(q, nq), sv = to_points(s, ns);
  std::vector<int> qbv(qb, qb + nb), sbv(sb, sb + nb);
  g_nb.clear();
  if (brute)
    batch_ordered_neighbors(qv, sv, qbv, sbv, g_nb, radius);
  else
    batch_nanoflann_neighbors(qv, sv, qbv, sbv, g_nb, radius);
  return nq > 0 ? static_cast<int>(g_nb.size() / static_cast<size_t>(nq)) : 0;
}

void ref_neighbors_fetch(int* out) {
  if (!g_nb.empty()) std::memcpy(out, g_nb.data(), g_nb.size() * sizeof(int));
}

// Returns the number of subsampled points; fetch with ref_subsample_fetch.
int ref_batch_grid_subsampling(const float* pts, int n, const int* lens, int nb, float dl, int max_p) {
  std::vector<PointXYZ> pv = to_points(pts, n);
  std::vector<int> lv(lens, lens + nb);
  std::vector<float> f0, f1;
  std::vector<int> c0, c1;
  g_sub.clear();
  g_sub_len.clear();
  batch_grid_subsampling(pv, g_sub, f0, f1, c0, c1, lv, g_sub_len, dl, max_p);
  return static_cast<int>(g_sub.size());
}

void ref_subsample_fetch(float* out_pts, int* out_lens) {
  for (size_t i = 0; i < g_sub.size(); ++i) {
    out_pts[3 * i] = g_sub[i].x;
    out_pts[3 * i + 1] = g_sub[i].y;
    out_pts[3 * i + 2] = g_sub[i].z;
  }
  for (size_t i = 0; i < g_sub_len.size(); ++i) out_lens[i] = g_sub_len[i];
}

}  // extern "C"
