// ORACLE - TEST INFRASTRUCTURE ONLY.
//
// Thin extern "C" driver around the REFERENCE's own sum-factorisation
// templates: `contract<T,Nk,Na,Nb,Nc,transpose>` and the 3-D
// `transpose<T,Na,Nb,Nc,offa,offb,offc>` are compiled from
// /root/reference/cpp/common/sum_factorisation.hpp where it lies (-I on the
// command line, see Makefile) - no reference source is copied into this repo.
// The rest of the reference's C++ operator (cpp/common/spectral_op.hpp)
// includes dolfinx/basix headers that do not exist in this image, so the cell
// loop of StiffnessSpectral3D::operator() (spectral_op.hpp:173-243) and
// MassSpectral3D::operator() (spectral_op.hpp:69-86) is restated here around
// those templates, with plain pointers where the reference has la::Vector.
//
// Output: oracle/_ref/libfus_ref.so (git-ignored, travels to the GPU box).

#include <array>
#include <cstdint>
#include <cstring>
#include <vector>

#include "sum_factorisation.hpp"  // the reference's file, not a copy

namespace {

template <typename T, int P>
void stiffness_cells(const T* x, const T* coeffs, T* y, const T* G_, const std::int32_t* dofmap,
                     const T* dphi_, std::int64_t c0, std::int64_t c1) {
  constexpr int N = P + 1;
  constexpr int Nd = N * N * N;
  std::array<T, Nd> x_, fw0_, fw1_, fw2_, y0_, y1_, y2_, T1, T2, T3, T4;
  T* fw0 = fw0_.data();
  T* fw1 = fw1_.data();
  T* fw2 = fw2_.data();

  for (std::int64_t c = c0; c < c1; ++c) {
    const std::int32_t* dm = dofmap + c * Nd;
    for (int i = 0; i < Nd; ++i) x_[i] = x[dm[i]];  // spectral_op.hpp:185-186

    T1.fill(0); T2.fill(0); T3.fill(0); T4.fill(0);

    fw0_.fill(0);  // :194-196
    contract<T, N, N, N, N, true>(dphi_, x_.data(), fw0);

    fw1_.fill(0);  // :199-203
    transpose<T, N, N, N, N, N * N, 1>(x_.data(), T1.data());
    contract<T, N, N, N, N, true>(dphi_, T1.data(), T2.data());
    transpose<T, N, N, N, N, N * N, 1>(T2.data(), fw1);

    fw2_.fill(0);  // :206-210
    transpose<T, N, N, N, 1, N, N * N>(x_.data(), T3.data());
    contract<T, N, N, N, N, true>(dphi_, T3.data(), T4.data());
    transpose<T, N, N, N, 1, N, N * N>(T4.data(), fw2);

    {  // stiffness::transform, spectral_op.hpp:113-130
      const T* G = G_ + c * Nd * 6;
      const T coeff = coeffs[c];
      for (int iq = 0; iq < Nd; ++iq) {
        const T* _G = G + iq * 6;
        const T w0 = fw0[iq], w1 = fw1[iq], w2 = fw2[iq];
        fw0[iq] = coeff * (_G[0] * w0 + _G[1] * w1 + _G[2] * w2);
        fw1[iq] = coeff * (_G[1] * w0 + _G[3] * w1 + _G[4] * w2);
        fw2[iq] = coeff * (_G[2] * w0 + _G[4] * w1 + _G[5] * w2);
      }
    }

    T1.fill(0); T2.fill(0); T3.fill(0); T4.fill(0);

    y0_.fill(0);  // :222-224
    contract<T, N, N, N, N, false>(dphi_, fw0, y0_.data());

    y1_.fill(0);  // :227-231
    transpose<T, N, N, N, N, N * N, 1>(fw1, T1.data());
    contract<T, N, N, N, N, false>(dphi_, T1.data(), T2.data());
    transpose<T, N, N, N, N, N * N, 1>(T2.data(), y1_.data());

    y2_.fill(0);  // :234-238
    transpose<T, N, N, N, 1, N, N * N>(fw2, T3.data());
    contract<T, N, N, N, N, false>(dphi_, T3.data(), T4.data());
    transpose<T, N, N, N, 1, N, N * N>(T4.data(), y2_.data());

    for (int i = 0; i < Nd; ++i) y[dm[i]] += y0_[i] + y1_[i] + y2_[i];  // :240-241
  }
}

template <typename T>
int stiffness_dispatch(int P, const T* x, const T* c, T* y, const T* G, const std::int32_t* dm,
                       const T* dphi, std::int64_t c0, std::int64_t c1) {
  switch (P) {
  case 2: stiffness_cells<T, 2>(x, c, y, G, dm, dphi, c0, c1); return 0;
  case 3: stiffness_cells<T, 3>(x, c, y, G, dm, dphi, c0, c1); return 0;
  case 4: stiffness_cells<T, 4>(x, c, y, G, dm, dphi, c0, c1); return 0;
  case 5: stiffness_cells<T, 5>(x, c, y, G, dm, dphi, c0, c1); return 0;
  case 6: stiffness_cells<T, 6>(x, c, y, G, dm, dphi, c0, c1); return 0;
  case 7: stiffness_cells<T, 7>(x, c, y, G, dm, dphi, c0, c1); return 0;
  }
  return 1;
}

// MassSpectral3D::operator(), spectral_op.hpp:69-86 (+ mass::transform :19-26)
template <typename T>
void mass_entities(const T* x, const T* coeffs, T* y, const T* detJ, const std::int32_t* dofmap,
                   std::int64_t e0, std::int64_t e1, int Nd) {
  std::vector<T> x_(Nd);
  for (std::int64_t c = e0; c < e1; ++c) {
    const std::int32_t* dm = dofmap + c * Nd;
    for (int i = 0; i < Nd; ++i) x_[i] = x[dm[i]];
    const T* sdetJ = detJ + c * Nd;
    for (int iq = 0; iq < Nd; ++iq) x_[iq] = coeffs[c] * x_[iq] * sdetJ[iq];
    for (int i = 0; i < Nd; ++i) y[dm[i]] += x_[i];
  }
}

template <typename T>
int stiffness_ranks(int P, const T* x, const T* c, T* y, std::int64_t ystride, const T* G,
                    const std::int32_t* dm, const T* dphi, std::int64_t ncells, int nranks) {
  int rc = 0;
  // one thread per emulated MPI rank: contiguous 1/k of the cells, private y
#pragma omp parallel for num_threads(nranks) schedule(static, 1) reduction(| : rc)
  for (int t = 0; t < nranks; ++t) {
    const std::int64_t a = ncells * t / nranks, b = ncells * (t + 1) / nranks;
    rc |= stiffness_dispatch<T>(P, x, c, y + t * ystride, G, dm, dphi, a, b);
  }
  return rc;
}

template <typename T>
void mass_ranks(const T* x, const T* c, T* y, std::int64_t ystride, const T* detJ,
                const std::int32_t* dm, std::int64_t nent, int ncols, int nranks) {
#pragma omp parallel for num_threads(nranks) schedule(static, 1)
  for (int t = 0; t < nranks; ++t) {
    const std::int64_t a = nent * t / nranks, b = nent * (t + 1) / nranks;
    mass_entities<T>(x, c, y + t * ystride, detJ, dm, a, b, ncols);
  }
}

}  // namespace

extern "C" {
int ref_stiffness_f64(const double* x, const double* c, double* y, const double* G,
                      const std::int32_t* dm, const double* dphi, std::int64_t ncells, int P) {
  return stiffness_dispatch<double>(P, x, c, y, G, dm, dphi, 0, ncells);
}
int ref_stiffness_f32(const float* x, const float* c, float* y, const float* G,
                      const std::int32_t* dm, const float* dphi, std::int64_t ncells, int P) {
  return stiffness_dispatch<float>(P, x, c, y, G, dm, dphi, 0, ncells);
}
int ref_stiffness_ranks_f64(const double* x, const double* c, double* y, std::int64_t ystride,
                            const double* G, const std::int32_t* dm, const double* dphi,
                            std::int64_t ncells, int P, int nranks) {
  return stiffness_ranks<double>(P, x, c, y, ystride, G, dm, dphi, ncells, nranks);
}
int ref_stiffness_ranks_f32(const float* x, const float* c, float* y, std::int64_t ystride,
                            const float* G, const std::int32_t* dm, const float* dphi,
                            std::int64_t ncells, int P, int nranks) {
  return stiffness_ranks<float>(P, x, c, y, ystride, G, dm, dphi, ncells, nranks);
}
void ref_mass_f64(const double* x, const double* c, double* y, const double* detJ,
                  const std::int32_t* dm, std::int64_t nent, int ncols) {
  mass_entities<double>(x, c, y, detJ, dm, 0, nent, ncols);
}
void ref_mass_f32(const float* x, const float* c, float* y, const float* detJ,
                  const std::int32_t* dm, std::int64_t nent, int ncols) {
  mass_entities<float>(x, c, y, detJ, dm, 0, nent, ncols);
}
void ref_mass_ranks_f64(const double* x, const double* c, double* y, std::int64_t ystride,
                        const double* detJ, const std::int32_t* dm, std::int64_t nent, int ncols,
                        int nranks) {
  mass_ranks<double>(x, c, y, ystride, detJ, dm, nent, ncols, nranks);
}
void ref_mass_ranks_f32(const float* x, const float* c, float* y, std::int64_t ystride,
                        const float* detJ, const std::int32_t* dm, std::int64_t nent, int ncols,
                        int nranks) {
  mass_ranks<float>(x, c, y, ystride, detJ, dm, nent, ncols, nranks);
}
}
