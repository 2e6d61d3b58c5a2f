// Header-only C++ host mirror of the slice of the TRIQS gf API that sits on the GPU hot path.
//
// Same names, argument meaning and error behaviour as the reference (TRIQS 3.3.1) for
//   gf<mesh, target>, block_gf                      c++/triqs/gfs/gf/gf.hpp:86-375, block/block_gf.hpp:104-295
//   mesh::imtime / imfreq / brzone / cyclat         c++/triqs/mesh/{imtime,imfreq,brzone,cyclat}.hpp
//   make_adjoint_mesh                               c++/triqs/mesh/adjoint.hpp:31-57
//   make_gf_from_fourier / fourier                  c++/triqs/gfs/transform/fourier.hpp:93-276
//   fit_tail / fit_hermitian_tail / make_zero_tail  c++/triqs/gfs/functions/functions2.hpp:50-162
//   inverse / invert_in_place                       c++/triqs/gfs/functions/functions2.hpp:197-209
//   g(tau) evaluation                               c++/triqs/gfs/evaluator.hpp:31-44
// but every numerical operation is a call into the C ABI of libtqgpu.so (include/tqgpu.h): this header contains
// no arithmetic of its own and there is no CPU path.  Data layout is the reference's: C-order
// [mesh..., target...] std::complex<double>, so arrays can be handed over from nda views without a copy.
#pragma once
#include <array>
#include <cmath>
#include <cstdio>
#include <complex>
#include <iostream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tqgpu.h"

namespace triqs_b200 {

  using dcomplex = std::complex<double>;

  // triqs::runtime_error  (c++/triqs/utility/exceptions.hpp:31-37)
  struct runtime_error : std::runtime_error {
    using std::runtime_error::runtime_error;
  };

  enum statistic_enum { Boson = 0, Fermion = 1 }; // c++/triqs/mesh/domains/matsubara.hpp

  // one context per thread (the reference is single threaded)
  inline tq_ctx *context() {
    struct holder {
      tq_ctx *c = nullptr;
      holder() {
        tq_status st = tq_create(0, &c);
        if (st != TQ_OK) throw runtime_error("triqs_b200: no CUDA device / tq_create failed (there is no CPU fallback)");
      }
      ~holder() { tq_destroy(c); }
    };
    thread_local holder h;
    return h.c;
  }

  inline void check(tq_status st) {
    if (st == TQ_OK) return;
    std::string msg = tq_last_error(context());
    throw runtime_error(msg.empty() ? "triqs_b200: status " + std::to_string(int(st)) : msg);
  }

  // the reference prints these to std::cerr (fourier_matsubara.cpp:109-110, 131-133, 208-210)
  inline void print_warnings(int bits) {
    if (bits & TQ_WARN_NOISY_TAU)
      std::cerr << "WARNING: Direct Fourier cannot deduce the high-frequency moments of G(tau) due to noise or a coarse tau-grid. "
                   "Please specify the high-frequency moments for higher accuracy.\n";
    if (bits & TQ_WARN_SHORT_TAU_MESH)
      std::cerr << "[Direct Fourier] WARNING: The imaginary time mesh is less than six times as long as the number of positive frequencies.\n"
                << "This can lead to substantial numerical inaccuracies at the boundary of the frequency mesh.\n";
    if (bits & TQ_WARN_TAIL_ERR)
      std::cerr << "WARNING: High frequency moments have an error greater than 1e-4.\n Please make sure you treat the constant offset analytically!\n";
  }

  // ------------------------------------------------------------------------------------------------
  // meshes
  // ------------------------------------------------------------------------------------------------
  namespace mesh {
    class imtime { // c++/triqs/mesh/imtime.hpp:36-122, bases/linear.hpp:47-55
      double _beta = 1;
      statistic_enum _stat = Fermion;
      long _n = 2;

      public:
      imtime() = default;
      imtime(double beta, statistic_enum s, long n_tau) : _beta(beta), _stat(s), _n(n_tau) {}
      double beta() const { return _beta; }
      statistic_enum statistic() const { return _stat; }
      long size() const { return _n; }
      double delta() const { return (_beta - 0.0) / (_n - 1); }
      double to_value(long i) const { // linear.hpp:154-160
        double wr = double(i) / (_n - 1);
        return 0.0 * (1 - wr) + _beta * wr;
      }
      double operator[](long i) const { return to_value(i); }
      bool operator==(imtime const &m) const { return _beta == m._beta && _stat == m._stat && _n == m._n; }
    };

    class imfreq { // c++/triqs/mesh/imfreq.hpp:43-304
      double _beta = 1;
      statistic_enum _stat = Fermion;
      long _n_iw = 0;

      public:
      imfreq() = default;
      imfreq(double beta, statistic_enum s, long n_iw = 1025) : _beta(beta), _stat(s), _n_iw(n_iw) {}
      double beta() const { return _beta; }
      statistic_enum statistic() const { return _stat; }
      long n_iw() const { return _n_iw; }
      long last_index() const { return _n_iw - 1; }
      long first_index() const { return -(last_index() + (_stat == Fermion ? 1 : 0)); }
      long size() const { return last_index() - first_index() + 1; }
      bool positive_only() const { return false; }
      dcomplex to_value(long n) const { return {0, M_PI * double(2 * n + _stat) / _beta}; } // matsubara.hpp:55
      dcomplex operator[](long data_index) const { return to_value(first_index() + data_index); }
      dcomplex w_max() const { return to_value(last_index()); }
      long to_data_index(long n) const { return n - first_index(); }
      bool operator==(imfreq const &m) const { return _beta == m._beta && _stat == m._stat && _n_iw == m._n_iw; }
    };

    struct lattice_mesh_base { // 3-index periodic meshes, data index i0*d1*d2 + i1*d2 + i2 (brzone.hpp:187-190)
      std::array<long, 3> _dims{1, 1, 1};
      std::array<long, 3> dims() const { return _dims; }
      long size() const { return _dims[0] * _dims[1] * _dims[2]; }
    };
    struct cyclat : lattice_mesh_base {
      cyclat() = default;
      cyclat(long d0, long d1 = 1, long d2 = 1) { _dims = {d0, d1, d2}; }
      bool operator==(cyclat const &m) const { return _dims == m._dims; }
    };
    struct brzone : lattice_mesh_base {
      brzone() = default;
      brzone(long d0, long d1 = 1, long d2 = 1) { _dims = {d0, d1, d2}; }
      bool operator==(brzone const &m) const { return _dims == m._dims; }
    };

    // n points on [xmin, xmax], both ends included (bases/linear.hpp:47-55, :154-160)
    struct linear_mesh_base {
      double _xmin = 0, _xmax = 0;
      long _n = 0;
      long size() const { return _n; }
      double delta() const { return _n == 1 ? 0. : (_xmax - _xmin) / (_n - 1); }
      double to_value(long i) const {
        if (_n == 1) return _xmin;
        double wr = double(i) / (_n - 1);
        return _xmin * (1 - wr) + _xmax * wr;
      }
      double operator[](long i) const { return to_value(i); }
    };
    struct retime : linear_mesh_base { // retime.hpp:39
      retime() = default;
      retime(double t_min, double t_max, long n_t) { _xmin = t_min, _xmax = t_max, _n = n_t; }
      double t_min() const { return _xmin; }
      double t_max() const { return _xmax; }
      bool operator==(retime const &m) const { return _xmin == m._xmin && _xmax == m._xmax && _n == m._n; }
    };
    struct refreq : linear_mesh_base { // refreq.hpp:38
      refreq() = default;
      refreq(double w_min, double w_max, long n_w) { _xmin = w_min, _xmax = w_max, _n = n_w; }
      double w_min() const { return _xmin; }
      double w_max() const { return _xmax; }
      bool operator==(refreq const &m) const { return _xmin == m._xmin && _xmax == m._xmax && _n == m._n; }
    };
    class legendre { // legendre.hpp: n_l coefficients of a gf on [0, beta]
      double _beta = 1;
      statistic_enum _stat = Fermion;
      long _n = 0;

      public:
      legendre() = default;
      legendre(double beta, statistic_enum s, long n_l) : _beta(beta), _stat(s), _n(n_l) {}
      double beta() const { return _beta; }
      statistic_enum statistic() const { return _stat; }
      long size() const { return _n; }
      bool operator==(legendre const &m) const { return _beta == m._beta && _stat == m._stat && _n == m._n; }
    };

    // c++/triqs/mesh/adjoint.hpp:31-57
    inline refreq make_adjoint_mesh(retime const &m, bool shift_half_bin = false) {
      long L = m.size();
      double wmax = M_PI * (L - 1) / (L * m.delta());
      if (shift_half_bin) return {-wmax + M_PI / L / m.delta(), wmax + M_PI / L / m.delta(), L};
      return {-wmax, wmax, L};
    }
    inline retime make_adjoint_mesh(refreq const &m, bool shift_half_bin = false) {
      long L = m.size();
      double tmax = M_PI * (L - 1) / (L * m.delta());
      if (shift_half_bin) return {-tmax + M_PI / L / m.delta(), tmax + M_PI / L / m.delta(), L};
      return {-tmax, tmax, L};
    }
    inline imfreq make_adjoint_mesh(imtime const &m, long n_iw = -1) {
      if (n_iw == -1) n_iw = (m.size() - 1) / 6;
      return {m.beta(), m.statistic(), n_iw};
    }
    inline imtime make_adjoint_mesh(imfreq const &m, long n_tau = -1) {
      if (n_tau == -1) n_tau = 6 * (m.last_index() + 1) + 1;
      return {m.beta(), m.statistic(), n_tau};
    }
    inline brzone make_adjoint_mesh(cyclat const &m) { return brzone{m._dims[0], m._dims[1], m._dims[2]}; }
    inline cyclat make_adjoint_mesh(brzone const &m) { return cyclat{m._dims[0], m._dims[1], m._dims[2]}; }

    // Cartesian product of two meshes (c++/triqs/mesh/prod.hpp:72-208): data index = i1 * size(m2) + i2
    template <typename M1, typename M2> struct prod {
      M1 m1;
      M2 m2;
      prod() = default;
      prod(M1 a, M2 b) : m1(std::move(a)), m2(std::move(b)) {}
      long size() const { return m1.size() * m2.size(); }
      bool operator==(prod const &o) const { return m1 == o.m1 && m2 == o.m2; }
    };
  } // namespace mesh

  // ------------------------------------------------------------------------------------------------
  // gf<mesh, target>
  // ------------------------------------------------------------------------------------------------
  struct scalar_valued {
    static constexpr int rank = 0;
    using scalar_t = dcomplex;
    using complex_t = scalar_valued;
  };
  struct matrix_valued {
    static constexpr int rank = 2;
    using scalar_t = dcomplex;
    using complex_t = matrix_valued;
  };
  template <int R> struct tensor_valued {
    static constexpr int rank = R;
    using scalar_t = dcomplex;
    using complex_t = tensor_valued<R>;
  };
  // real-valued targets (c++/triqs/gfs/gf/targets.hpp: scalar_real_valued, matrix_real_valued, tensor_real_valued<R>): double data;
  // the Fourier transform of such a gf promotes it to T::complex_t (fourier.hpp:66, :78-81)
  struct scalar_real_valued {
    static constexpr int rank = 0;
    using scalar_t = double;
    using complex_t = scalar_valued;
  };
  struct matrix_real_valued {
    static constexpr int rank = 2;
    using scalar_t = double;
    using complex_t = matrix_valued;
  };
  template <int R> struct tensor_real_valued {
    static constexpr int rank = R;
    using scalar_t = double;
    using complex_t = tensor_valued<R>;
  };
  namespace detail {
    template <typename T> struct real_of;
    template <> struct real_of<scalar_valued> { using type = scalar_real_valued; };
    template <> struct real_of<matrix_valued> { using type = matrix_real_valued; };
    template <int R> struct real_of<tensor_valued<R>> { using type = tensor_real_valued<R>; };
    template <typename T> constexpr bool is_real_target = std::is_same<typename T::scalar_t, double>::value;
  } // namespace detail

  // moments / small dense arrays: C-order [n_rows, target...] flattened to [n_rows][n_others]
  struct moment_array {
    long n_rows = 0, n_others = 0;
    std::vector<dcomplex> v;
    moment_array() = default;
    moment_array(long r, long c) : n_rows(r), n_others(c), v(size_t(r * c), dcomplex{0, 0}) {}
    dcomplex &operator()(long r, long c) { return v[size_t(r * n_others + c)]; }
    dcomplex const &operator()(long r, long c) const { return v[size_t(r * n_others + c)]; }
    bool is_empty() const { return v.empty(); }
  };

  template <typename Mesh, typename Target = matrix_valued> class gf {
    public:
    using mesh_t = Mesh;
    using target_t = Target;
    using scalar_t = typename Target::scalar_t; // dcomplex, or double for the *_real_valued targets

    private:
    Mesh _mesh;
    std::vector<long> _tshape; // target shape (empty for scalar_valued)
    std::vector<scalar_t> _data;

    public:
    gf() = default;
    gf(Mesh m, std::vector<long> target_shape = {}) : _mesh(std::move(m)), _tshape(std::move(target_shape)) {
      if ((int)_tshape.size() != Target::rank) throw runtime_error("gf: target shape does not match the target rank");
      _data.assign(size_t(_mesh.size() * n_others()), scalar_t{});
    }
    Mesh const &mesh() const { return _mesh; }
    std::vector<long> const &target_shape() const { return _tshape; }
    long n_others() const {
      long n = 1;
      for (long s : _tshape) n *= s;
      return n;
    }
    // inner matrix dimension for the hermitian fit (functions2.hpp:96-106)
    long inner_matrix_dim() const {
      if (Target::rank == 0) return 1;
      if (Target::rank == 2 && _tshape[0] == _tshape[1]) return _tshape[0];
      throw runtime_error("Incompatible target_shape for fit_hermitian_tail");
    }
    std::vector<scalar_t> &data() { return _data; }
    std::vector<scalar_t> const &data() const { return _data; }
    // value of the target at mesh data index i (pointer to n_others contiguous entries, C order)
    scalar_t *operator[](long i) { return _data.data() + size_t(i * n_others()); }
    scalar_t const *operator[](long i) const { return _data.data() + size_t(i * n_others()); }
    // off-mesh evaluation g(tau): linear interpolation on the device (evaluator.hpp:35-43)
    std::vector<dcomplex> operator()(double x) const;
  };

  template <typename Mesh, typename Target = matrix_valued> class block_gf { // block/block_gf.hpp:104-295
    std::vector<std::string> _names;
    std::vector<gf<Mesh, Target>> _blocks;

    public:
    block_gf() = default;
    block_gf(std::vector<std::string> names, std::vector<gf<Mesh, Target>> blocks) : _names(std::move(names)), _blocks(std::move(blocks)) {
      if (_names.size() != _blocks.size()) throw runtime_error("block_gf: names and blocks differ in size");
    }
    long size() const { return (long)_blocks.size(); }
    std::vector<std::string> const &block_names() const { return _names; }
    gf<Mesh, Target> &operator[](long i) { return _blocks[size_t(i)]; }
    gf<Mesh, Target> const &operator[](long i) const { return _blocks[size_t(i)]; }
  };

  template <typename G> moment_array make_zero_tail(G const &g, int n_moments = 10) { // functions2.hpp:154-162
    return moment_array(n_moments, g.n_others());
  }

  // ------------------------------------------------------------------------------------------------
  // Fourier  (fourier.hpp:93-276)
  // ------------------------------------------------------------------------------------------------
  namespace detail {
    inline const tq_complex *cptr(std::vector<dcomplex> const &v) { return reinterpret_cast<const tq_complex *>(v.data()); }
    inline tq_complex *ptr(std::vector<dcomplex> &v) { return reinterpret_cast<tq_complex *>(v.data()); }

    // same-shape blocks go to the device as ONE batched call; otherwise block by block like map_block_gf (block/map.hpp:35-40)
    template <typename G> bool same_shape(std::vector<G const *> const &gs) {
      for (auto *g : gs)
        if (!(g->mesh() == gs[0]->mesh()) || g->target_shape() != gs[0]->target_shape()) return false;
      return true;
    }
  } // namespace detail

  // g_out() = fourier(g_in[, known_moments]) on an existing gf (the lazy assignment of fourier.hpp:257-276)
  template <typename T>
  void set_from_fourier(gf<mesh::imfreq, T> &gw, gf<mesh::imtime, T> const &gt, moment_array const &known_moments = {}) {
    if (gw.target_shape() != gt.target_shape()) throw runtime_error("fourier: target shapes differ");
    int warn = 0;
    check(tq_fourier_tau_to_iw(context(), gt.mesh().beta(), gt.mesh().statistic(), gt.mesh().size(), gw.mesh().n_iw(), gt.n_others(), 1,
                               detail::cptr(gt.data()), known_moments.is_empty() ? nullptr : detail::cptr(known_moments.v),
                               (int)known_moments.n_rows, detail::ptr(gw.data()), &warn));
    print_warnings(warn);
  }
  template <typename T>
  void set_from_fourier(gf<mesh::imtime, T> &gt, gf<mesh::imfreq, T> const &gw, moment_array const &known_moments = {}) {
    if (gw.target_shape() != gt.target_shape()) throw runtime_error("fourier: target shapes differ");
    int warn = 0;
    check(tq_fourier_iw_to_tau(context(), gw.mesh().beta(), gw.mesh().statistic(), gw.mesh().n_iw(), gt.mesh().size(), gw.n_others(), 1,
                               detail::cptr(gw.data()), known_moments.is_empty() ? nullptr : detail::cptr(known_moments.v),
                               (int)known_moments.n_rows, detail::ptr(gt.data()), &warn, nullptr));
    print_warnings(warn);
  }
  // real-valued G(tau): only the doubles go to the device, promoted there (fourier.hpp:66 static_assert T1::complex_t == T2, :78-81)
  template <typename T, typename = std::enable_if_t<detail::is_real_target<T>>>
  void set_from_fourier(gf<mesh::imfreq, typename T::complex_t> &gw, gf<mesh::imtime, T> const &gt, moment_array const &known_moments = {}) {
    if (gw.target_shape() != gt.target_shape()) throw runtime_error("fourier: target shapes differ");
    int warn = 0;
    check(tq_fourier_tau_to_iw_real(context(), gt.mesh().beta(), gt.mesh().statistic(), gt.mesh().size(), gw.mesh().n_iw(), gt.n_others(), 1,
                                    gt.data().data(), known_moments.is_empty() ? nullptr : detail::cptr(known_moments.v),
                                    (int)known_moments.n_rows, detail::ptr(gw.data()), &warn));
    print_warnings(warn);
  }
  // is_gf_real / real(g)  (functions2.hpp:223-246)
  template <typename M, typename T> bool is_gf_real(gf<M, T> const &g, double tolerance = 1.e-13) {
    if constexpr (detail::is_real_target<T>) {
      return true;
    } else {
      for (auto const &z : g.data())
        if (std::abs(z.imag()) > tolerance) return false;
      return true;
    }
  }
  template <typename M, typename T> gf<M, typename detail::real_of<T>::type> real(gf<M, T> const &g) {
    gf<M, typename detail::real_of<T>::type> r{g.mesh(), g.target_shape()};
    for (size_t i = 0; i < g.data().size(); ++i) r.data()[i] = g.data()[i].real();
    return r;
  }
  // make_gf_from_fourier(gw) + is_gf_real + real(g) in one call: only Re G(tau) comes back from the device; throws if
  // max |Im G(tau)| exceeds the tolerance
  template <typename T>
  gf<mesh::imtime, typename detail::real_of<T>::type> make_real_gf_from_fourier(gf<mesh::imfreq, T> const &gw, long n_tau = -1,
                                                                                 moment_array const &known_moments = {}, double tolerance = 1.e-13) {
    gf<mesh::imtime, typename detail::real_of<T>::type> gt{mesh::make_adjoint_mesh(gw.mesh(), n_tau), gw.target_shape()};
    int warn = 0;
    double max_imag = 0;
    check(tq_fourier_iw_to_tau_real(context(), gw.mesh().beta(), gw.mesh().statistic(), gw.mesh().n_iw(), gt.mesh().size(), gw.n_others(), 1,
                                    detail::cptr(gw.data()), known_moments.is_empty() ? nullptr : detail::cptr(known_moments.v),
                                    (int)known_moments.n_rows, gt.data().data(), &warn, nullptr, &max_imag));
    print_warnings(warn);
    if (max_imag > tolerance) {
      char buf[64];
      std::snprintf(buf, sizeof buf, "%.3e", max_imag);
      throw runtime_error(std::string("make_real_gf_from_fourier: G(tau) is not real, max |Im| = ") + buf);
    }
    return gt;
  }
  template <typename T> void set_from_fourier(gf<mesh::brzone, T> &gk, gf<mesh::cyclat, T> const &gr) {
    auto d = gr.mesh().dims();
    long dims[3] = {d[0], d[1], d[2]};
    check(tq_fourier_lattice(context(), dims, 1, gr.n_others(), detail::cptr(gr.data()), detail::ptr(gk.data())));
  }
  template <typename T> void set_from_fourier(gf<mesh::cyclat, T> &gr, gf<mesh::brzone, T> const &gk) {
    auto d = gk.mesh().dims();
    long dims[3] = {d[0], d[1], d[2]};
    check(tq_fourier_lattice(context(), dims, 0, gk.n_others(), detail::cptr(gk.data()), detail::ptr(gr.data())));
  }

  // retime <-> refreq  (fourier_real.cpp:35-131)
  template <typename T>
  void set_from_fourier(gf<mesh::refreq, T> &gw, gf<mesh::retime, T> const &gt, moment_array const &known_moments = {}) {
    if (gw.target_shape() != gt.target_shape()) throw runtime_error("fourier: target shapes differ");
    if (gw.mesh().size() != gt.mesh().size()) throw runtime_error("Meshes are different");
    check(tq_fourier_t_to_w(context(), gt.mesh().t_min(), gt.mesh().t_max(), gw.mesh().w_min(), gw.mesh().w_max(), gt.mesh().size(), gt.n_others(),
                            1, detail::cptr(gt.data()), known_moments.is_empty() ? nullptr : detail::cptr(known_moments.v),
                            (int)known_moments.n_rows, detail::ptr(gw.data())));
  }
  template <typename T>
  void set_from_fourier(gf<mesh::retime, T> &gt, gf<mesh::refreq, T> const &gw, moment_array const &known_moments = {}) {
    if (gw.target_shape() != gt.target_shape()) throw runtime_error("fourier: target shapes differ");
    if (gw.mesh().size() != gt.mesh().size()) throw runtime_error("Meshes are of different size");
    check(tq_fourier_w_to_t(context(), gt.mesh().t_min(), gt.mesh().t_max(), gw.mesh().w_min(), gw.mesh().w_max(), gw.mesh().size(), gw.n_others(),
                            1, detail::cptr(gw.data()), known_moments.is_empty() ? nullptr : detail::cptr(known_moments.v),
                            (int)known_moments.n_rows, detail::ptr(gt.data())));
  }
  template <typename T>
  gf<mesh::retime, T> make_gf_from_fourier(gf<mesh::refreq, T> const &gw, mesh::retime const &m, moment_array const &known_moments = {}) {
    gf<mesh::retime, T> gt{m, gw.target_shape()};
    set_from_fourier(gt, gw, known_moments);
    return gt;
  }
  template <typename T>
  gf<mesh::refreq, T> make_gf_from_fourier(gf<mesh::retime, T> const &gt, mesh::refreq const &m, moment_array const &known_moments = {}) {
    gf<mesh::refreq, T> gw{m, gt.target_shape()};
    set_from_fourier(gw, gt, known_moments);
    return gw;
  }
  template <typename T> gf<mesh::retime, T> make_gf_from_fourier(gf<mesh::refreq, T> const &gw, bool shift_half_bin = false) {
    return make_gf_from_fourier(gw, mesh::make_adjoint_mesh(gw.mesh(), shift_half_bin));
  }
  template <typename T> gf<mesh::refreq, T> make_gf_from_fourier(gf<mesh::retime, T> const &gt, bool shift_half_bin = false) {
    return make_gf_from_fourier(gt, mesh::make_adjoint_mesh(gt.mesh(), shift_half_bin));
  }

  // Legendre <-> Matsubara  (legendre_matsubara.hpp:37-118)
  template <typename T> void legendre_matsubara_direct(gf<mesh::imfreq, T> &gw, gf<mesh::legendre, T> const &gl) {
    check(tq_legendre_to_imfreq(context(), gw.mesh().statistic(), gl.mesh().size(), gw.mesh().n_iw(), gl.n_others(), 1, detail::cptr(gl.data()),
                                detail::ptr(gw.data())));
  }
  template <typename T> void legendre_matsubara_direct(gf<mesh::imtime, T> &gt, gf<mesh::legendre, T> const &gl) {
    check(tq_legendre_to_imtime(context(), gt.mesh().beta(), gl.mesh().size(), gt.mesh().size(), gl.n_others(), 1, detail::cptr(gl.data()),
                                detail::ptr(gt.data())));
  }
  template <typename T> void legendre_matsubara_inverse(gf<mesh::legendre, T> &gl, gf<mesh::imfreq, T> const &gw) {
    check(tq_imfreq_to_legendre(context(), gw.mesh().beta(), gw.mesh().statistic(), gw.mesh().n_iw(), gl.mesh().size(), gw.n_others(), 1,
                                detail::cptr(gw.data()), detail::ptr(gl.data())));
  }
  template <typename T> void legendre_matsubara_inverse(gf<mesh::legendre, T> &gl, gf<mesh::imtime, T> const &gt) {
    check(tq_imtime_to_legendre(context(), gt.mesh().beta(), gt.mesh().size(), gl.mesh().size(), gt.n_others(), 1, detail::cptr(gt.data()),
                                detail::ptr(gl.data())));
  }

  // factories  (fourier.hpp:122-144, 238-245)
  template <typename T>
  gf<mesh::imfreq, typename T::complex_t> make_gf_from_fourier(gf<mesh::imtime, T> const &gt, long n_iw = -1, moment_array const &known_moments = {}) {
    gf<mesh::imfreq, typename T::complex_t> gw{mesh::make_adjoint_mesh(gt.mesh(), n_iw), gt.target_shape()};
    set_from_fourier(gw, gt, known_moments);
    return gw;
  }
  template <typename T>
  gf<mesh::imtime, T> make_gf_from_fourier(gf<mesh::imfreq, T> const &gw, long n_tau = -1, moment_array const &known_moments = {}) {
    gf<mesh::imtime, T> gt{mesh::make_adjoint_mesh(gw.mesh(), n_tau), gw.target_shape()};
    set_from_fourier(gt, gw, known_moments);
    return gt;
  }
  template <typename T> gf<mesh::brzone, T> make_gf_from_fourier(gf<mesh::cyclat, T> const &gr) {
    gf<mesh::brzone, T> gk{mesh::make_adjoint_mesh(gr.mesh()), gr.target_shape()};
    set_from_fourier(gk, gr);
    return gk;
  }
  template <typename T> gf<mesh::cyclat, T> make_gf_from_fourier(gf<mesh::brzone, T> const &gk) {
    gf<mesh::cyclat, T> gr{mesh::make_adjoint_mesh(gk.mesh()), gk.target_shape()};
    set_from_fourier(gr, gk);
    return gr;
  }

  // block_gf: all blocks of equal shape are transformed by one batched launch  (fourier.hpp:189-230)
  template <typename T> block_gf<mesh::imfreq, T> make_gf_from_fourier(block_gf<mesh::imtime, T> const &bgt, long n_iw = -1) {
    std::vector<gf<mesh::imfreq, T>> out;
    std::vector<gf<mesh::imtime, T> const *> gs;
    for (long b = 0; b < bgt.size(); ++b) gs.push_back(&bgt[b]);
    if (bgt.size() > 1 && detail::same_shape(gs)) {
      auto const &g0 = bgt[0];
      auto mw = mesh::make_adjoint_mesh(g0.mesh(), n_iw);
      long nb = bgt.size(), no = g0.n_others();
      std::vector<dcomplex> in(size_t(nb * g0.mesh().size() * no)), res(size_t(nb * mw.size() * no));
      for (long b = 0; b < nb; ++b) std::copy(bgt[b].data().begin(), bgt[b].data().end(), in.begin() + size_t(b * g0.mesh().size() * no));
      std::vector<int> warn(size_t(nb), 0);
      check(tq_fourier_tau_to_iw(context(), g0.mesh().beta(), g0.mesh().statistic(), g0.mesh().size(), mw.n_iw(), no, nb, detail::cptr(in),
                                 nullptr, 0, detail::ptr(res), warn.data()));
      int w = 0;
      for (int x : warn) w |= x;
      print_warnings(w);
      for (long b = 0; b < nb; ++b) {
        gf<mesh::imfreq, T> g{mw, g0.target_shape()};
        std::copy(res.begin() + size_t(b * mw.size() * no), res.begin() + size_t((b + 1) * mw.size() * no), g.data().begin());
        out.push_back(std::move(g));
      }
    } else {
      for (long b = 0; b < bgt.size(); ++b) out.push_back(make_gf_from_fourier(bgt[b], n_iw));
    }
    return {bgt.block_names(), std::move(out)};
  }
  template <typename T> block_gf<mesh::imtime, T> make_gf_from_fourier(block_gf<mesh::imfreq, T> const &bgw, long n_tau = -1) {
    std::vector<gf<mesh::imtime, T>> out;
    std::vector<gf<mesh::imfreq, T> const *> gs;
    for (long b = 0; b < bgw.size(); ++b) gs.push_back(&bgw[b]);
    if (bgw.size() > 1 && detail::same_shape(gs)) {
      auto const &g0 = bgw[0];
      auto mt = mesh::make_adjoint_mesh(g0.mesh(), n_tau);
      long nb = bgw.size(), no = g0.n_others();
      std::vector<dcomplex> in(size_t(nb * g0.mesh().size() * no)), res(size_t(nb * mt.size() * no));
      for (long b = 0; b < nb; ++b) std::copy(bgw[b].data().begin(), bgw[b].data().end(), in.begin() + size_t(b * g0.mesh().size() * no));
      std::vector<int> warn(size_t(nb), 0);
      check(tq_fourier_iw_to_tau(context(), g0.mesh().beta(), g0.mesh().statistic(), g0.mesh().n_iw(), mt.size(), no, nb, detail::cptr(in),
                                 nullptr, 0, detail::ptr(res), warn.data(), nullptr));
      int w = 0;
      for (int x : warn) w |= x;
      print_warnings(w);
      for (long b = 0; b < nb; ++b) {
        gf<mesh::imtime, T> g{mt, g0.target_shape()};
        std::copy(res.begin() + size_t(b * mt.size() * no), res.begin() + size_t((b + 1) * mt.size() * no), g.data().begin());
        out.push_back(std::move(g));
      }
    } else {
      for (long b = 0; b < bgw.size(); ++b) out.push_back(make_gf_from_fourier(bgw[b], n_tau));
    }
    return {bgw.block_names(), std::move(out)};
  }

  // ------------------------------------------------------------------------------------------------
  // tail fits  (functions2.hpp:50-134)
  // ------------------------------------------------------------------------------------------------
  struct tail_fit_parameters { // tail_fitter.hpp:302-321 (mesh.set_tail_fit_parameters)
    double tail_fraction = 0.2;
    int n_tail_max = 30;
    int expansion_order = -1; // adaptive
  };

  namespace detail {
    template <typename T>
    std::pair<moment_array, double> fit_impl(gf<mesh::imfreq, T> const &g, moment_array const &known, bool herm, tail_fit_parameters p) {
      if (!known.is_empty() && known.n_others != g.n_others()) throw runtime_error("known_moments shape incompatible with shape of data");
      moment_array out(std::max<long>(TQ_MAX_MOMENTS, known.n_rows), g.n_others());
      int nm = 0;
      double err = 0;
      check(tq_fit_tail(context(), g.mesh().beta(), g.mesh().statistic(), g.mesh().n_iw(), p.tail_fraction, p.n_tail_max, p.expansion_order,
                        herm ? 1 : 0, herm ? (int)g.inner_matrix_dim() : 1, g.n_others(), 1, cptr(g.data()),
                        known.is_empty() ? nullptr : cptr(known.v), (int)known.n_rows, ptr(out.v), &nm, &err));
      out.n_rows = nm;
      out.v.resize(size_t(nm * g.n_others()));
      return {std::move(out), err};
    }
  } // namespace detail

  template <typename T>
  std::pair<moment_array, double> fit_tail(gf<mesh::imfreq, T> const &g, moment_array const &known_moments = {}, tail_fit_parameters p = {}) {
    return detail::fit_impl(g, known_moments, false, p);
  }
  template <typename T>
  std::pair<moment_array, double> fit_hermitian_tail(gf<mesh::imfreq, T> const &g, moment_array const &known_moments = {},
                                                     tail_fit_parameters p = {}) {
    return detail::fit_impl(g, known_moments, true, p);
  }
  // block overloads return the maximum error (functions2.hpp:66-78, 121-134)
  template <typename T>
  std::pair<std::vector<moment_array>, double> fit_hermitian_tail(block_gf<mesh::imfreq, T> const &bg, tail_fit_parameters p = {}) {
    std::vector<moment_array> tails;
    double max_err = 0;
    for (long b = 0; b < bg.size(); ++b) {
      auto [t, e] = fit_hermitian_tail(bg[b], {}, p);
      tails.push_back(std::move(t));
      max_err = std::max(max_err, e);
    }
    return {std::move(tails), max_err};
  }

  // ------------------------------------------------------------------------------------------------
  // density  (c++/triqs/gfs/functions/density.cpp:32-128): n x n matrix of a matrix-valued gf, the scalar for a scalar-valued one
  // ------------------------------------------------------------------------------------------------
  inline std::vector<dcomplex> density(gf<mesh::imfreq, matrix_valued> const &g, moment_array const &known_moments = {}) {
    auto const &ts = g.target_shape();
    if (ts[0] != ts[1]) throw runtime_error("density: target is not square");
    std::vector<dcomplex> res(size_t(ts[0] * ts[1]));
    check(tq_density(context(), g.mesh().beta(), g.mesh().statistic(), g.mesh().n_iw(), (int)ts[0], 1, detail::cptr(g.data()),
                     known_moments.is_empty() ? nullptr : detail::cptr(known_moments.v), (int)known_moments.n_rows, detail::ptr(res), nullptr,
                     nullptr));
    return res;
  }
  inline dcomplex density(gf<mesh::imfreq, scalar_valued> const &g, moment_array const &known_moments = {}) {
    std::vector<dcomplex> res(1);
    check(tq_density(context(), g.mesh().beta(), g.mesh().statistic(), g.mesh().n_iw(), 1, 1, detail::cptr(g.data()),
                     known_moments.is_empty() ? nullptr : detail::cptr(known_moments.v), (int)known_moments.n_rows, detail::ptr(res), nullptr,
                     nullptr));
    return res[0];
  }

  // ------------------------------------------------------------------------------------------------
  // hermiticity helpers, replace_by_tail  (gfs/functions/imfreq.hpp:81-124, :175-215, :258-267)
  // ------------------------------------------------------------------------------------------------
  namespace detail {
    template <typename M> constexpr int mesh_kind() { return std::is_same<M, mesh::imtime>::value ? 1 : 0; }
    template <typename G> long herm_dim(G const &g) { // matrix: n;  rank-4 tensor [a,b,c,d]: the (ab),(cd) matrix
      auto const &ts = g.target_shape();
      if (ts.empty()) return 1;
      if (ts.size() == 2 && ts[0] == ts[1]) return ts[0];
      if (ts.size() == 4 && ts[0] * ts[1] == ts[2] * ts[3]) return ts[0] * ts[1];
      throw runtime_error("hermiticity: incompatible target shape");
    }
  } // namespace detail
  template <typename M, typename T> bool is_gf_hermitian(gf<M, T> const &g, double tolerance = 1.e-12) {
    int res = 0;
    check(tq_is_gf_hermitian(context(), detail::mesh_kind<M>(), g.mesh().size(), (int)detail::herm_dim(g), 1, detail::cptr(g.data()), tolerance,
                             &res));
    return res != 0;
  }
  template <typename M, typename T> gf<M, T> make_hermitian(gf<M, T> const &g) {
    gf<M, T> out{g.mesh(), g.target_shape()};
    check(tq_make_hermitian(context(), detail::mesh_kind<M>(), g.mesh().size(), (int)detail::herm_dim(g), 1, detail::cptr(g.data()),
                            detail::ptr(out.data())));
    return out;
  }
  template <typename T> void replace_by_tail(gf<mesh::imfreq, T> &g, moment_array const &tail, int n_min) {
    check(tq_replace_by_tail(context(), g.mesh().beta(), g.mesh().statistic(), g.mesh().n_iw(), g.n_others(), 1, detail::ptr(g.data()),
                             detail::cptr(tail.v), (int)tail.n_rows, n_min));
  }

  // eps_k on a brzone mesh from hoppings  (lattice/tight_binding.hpp:88-123)
  struct tight_binding {
    std::vector<std::array<int, 3>> displ_vec;
    std::vector<std::vector<dcomplex>> overlap_mat_vec; // each [n_orb * n_orb]
    int n_orb = 1;
    gf<mesh::brzone, matrix_valued> fourier(mesh::brzone const &k_mesh) const {
      gf<mesh::brzone, matrix_valued> out{k_mesh, {n_orb, n_orb}};
      std::vector<int> d;
      std::vector<dcomplex> h;
      for (auto const &r : displ_vec) d.insert(d.end(), r.begin(), r.end());
      for (auto const &m : overlap_mat_vec) h.insert(h.end(), m.begin(), m.end());
      auto dm = k_mesh.dims();
      long dims[3] = {dm[0], dm[1], dm[2]};
      check(tq_tight_binding_fourier(context(), n_orb, (int)displ_vec.size(), d.data(), detail::cptr(h), dims, detail::ptr(out.data())));
      return out;
    }
  };

  // ------------------------------------------------------------------------------------------------
  // inverse  (functions2.hpp:197-209)
  // ------------------------------------------------------------------------------------------------
  template <typename M> void invert_in_place(gf<M, matrix_valued> &g) {
    auto const &ts = g.target_shape();
    if (ts[0] != ts[1]) throw runtime_error("invert_in_place: target is not square");
    check(tq_invert_batched(context(), (int)ts[0], g.mesh().size(), detail::ptr(g.data())));
  }
  template <typename M> gf<M, matrix_valued> inverse(gf<M, matrix_valued> const &g) {
    auto r = g;
    invert_in_place(r);
    return r;
  }

  // G_loc(iw) = sum_k w_k [(iw + mu) - eps_k - Sigma(iw)]^-1   (python/triqs/sumk/sumk_discrete.py:140-166)
  inline gf<mesh::imfreq, matrix_valued> dyson_k_sum(gf<mesh::brzone, matrix_valued> const &eps_k, gf<mesh::imfreq, matrix_valued> const &sigma,
                                                     double mu, bool all_reduce = false) {
    gf<mesh::imfreq, matrix_valued> out{sigma.mesh(), sigma.target_shape()};
    long n = sigma.target_shape()[0];
    check(tq_dyson_ksum(context(), (int)n, eps_k.mesh().size(), sigma.mesh().n_iw(), sigma.mesh().beta(), sigma.mesh().statistic(), mu,
                        detail::cptr(eps_k.data()), nullptr, 1.0 / double(eps_k.mesh().size()), detail::cptr(sigma.data()),
                        detail::ptr(out.data()), all_reduce ? 1 : 0));
    return out;
  }

  // ------------------------------------------------------------------------------------------------
  // evaluation  (evaluator.hpp:35-43; batched form for many points)
  // ------------------------------------------------------------------------------------------------
  template <typename T> std::vector<dcomplex> evaluate(gf<mesh::imtime, T> const &g, std::vector<double> const &taus) {
    std::vector<dcomplex> out(size_t(taus.size() * g.n_others()));
    check(tq_interp_tau(context(), g.mesh().beta(), g.mesh().size(), g.n_others(), detail::cptr(g.data()), (long)taus.size(), taus.data(),
                        detail::ptr(out)));
    return out;
  }
  // gf_evaluator<imfreq> (evaluator.hpp:47-75), batched over Matsubara indices: the mesh value inside the window, the fitted
  // high-frequency tail outside.  Returns [ns.size()][n_others].
  template <typename T> std::vector<dcomplex> evaluate(gf<mesh::imfreq, T> const &g, std::vector<long> const &ns) {
    std::vector<dcomplex> out(size_t(ns.size() * g.n_others()));
    double err = 0;
    check(tq_evaluate_imfreq(context(), g.mesh().beta(), g.mesh().statistic(), g.mesh().n_iw(), g.n_others(), 1, detail::cptr(g.data()),
                             (long)ns.size(), ns.data(), detail::ptr(out), &err));
    return out;
  }
  // Product-mesh evaluation  g(x1, x2)  (c++/triqs/mesh/evaluate.hpp:97-114), batched over points: multi-linear interpolation with
  // the per-component rules of linear.hpp:221-228 (imtime / retime / refreq: the value), brzone.hpp:355-373 (a k vector given in units
  // of the mesh index, i.e. transpose(units_inv) * k) and index pass-through (imfreq: the Matsubara index n, inside the mesh).
  namespace detail {
    inline void push_axes(mesh::imtime const &m, std::vector<tq_axis> &a) { a.push_back(tq_axis{TQ_AXIS_LINEAR, m.size(), 0.0, m.beta()}); }
    inline void push_axes(mesh::linear_mesh_base const &m, std::vector<tq_axis> &a) { a.push_back(tq_axis{TQ_AXIS_LINEAR, m.size(), m._xmin, m._xmax}); }
    inline void push_axes(mesh::brzone const &m, std::vector<tq_axis> &a) {
      for (int j = 0; j < 3; ++j) a.push_back(tq_axis{TQ_AXIS_PERIODIC, m._dims[j], 0.0, 0.0});
    }
    inline void push_axes(mesh::imfreq const &m, std::vector<tq_axis> &a) { a.push_back(tq_axis{TQ_AXIS_INDEX, m.size(), 0.0, 0.0}); }
    inline void push_coord(mesh::imtime const &, double x, std::vector<double> &c) { c.push_back(x); }
    inline void push_coord(mesh::linear_mesh_base const &, double x, std::vector<double> &c) { c.push_back(x); }
    inline void push_coord(mesh::brzone const &, std::array<double, 3> const &k, std::vector<double> &c) { c.insert(c.end(), k.begin(), k.end()); }
    inline void push_coord(mesh::imfreq const &m, long n, std::vector<double> &c) {
      if (n < m.first_index() || n > m.last_index()) throw runtime_error("product-mesh evaluation: Matsubara index outside the mesh");
      c.push_back(double(n - m.first_index()));
    }
  } // namespace detail
  template <typename M1, typename M2, typename T, typename X1, typename X2>
  std::vector<dcomplex> evaluate(gf<mesh::prod<M1, M2>, T> const &g, std::vector<X1> const &x1, std::vector<X2> const &x2) {
    if (x1.size() != x2.size()) throw runtime_error("product-mesh evaluation: one coordinate per component and point");
    std::vector<tq_axis> axes;
    detail::push_axes(g.mesh().m1, axes);
    detail::push_axes(g.mesh().m2, axes);
    std::vector<double> coords;
    for (size_t i = 0; i < x1.size(); ++i) {
      detail::push_coord(g.mesh().m1, x1[i], coords);
      detail::push_coord(g.mesh().m2, x2[i], coords);
    }
    std::vector<dcomplex> out(size_t(x1.size() * g.n_others()));
    check(tq_evaluate_prod(context(), (int)axes.size(), axes.data(), g.n_others(), detail::cptr(g.data()), (long)x1.size(), coords.data(),
                           detail::ptr(out)));
    return out;
  }
  // engine options of the context (tq_set_option: never change results), e.g. set_option("host_lanes", 4)
  inline void set_option(const char *name, long value) { check(tq_set_option(context(), name, value)); }
  // which engine / plan an imtime <-> imfreq transform of this shape takes, as JSON (host only: needs no device) -- tq_plan_describe
  inline std::string plan_describe(mesh::imtime const &mt, long n_iw, long n_others = 1, long batch = 1) {
    char buf[1024];
    if (tq_plan_describe(mt.size(), n_iw, mt.statistic(), n_others, batch, buf, sizeof buf) != TQ_OK) throw runtime_error("plan_describe: invalid argument");
    return buf;
  }

  template <typename Mesh, typename Target> std::vector<dcomplex> gf<Mesh, Target>::operator()(double x) const {
    static_assert(std::is_same<Mesh, mesh::imtime>::value, "g(x) is implemented for imtime gfs");
    return evaluate(*this, std::vector<double>{x});
  }

} // namespace triqs_b200
