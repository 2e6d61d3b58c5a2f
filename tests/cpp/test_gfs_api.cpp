// C++ host-side test of the drop-in API (triqs_b200/cpp/triqs_b200/gfs.hpp) -- reads like the reference's own tests:
//   test/c++/gfs/fourier_matsubara.cpp:23-114, fourier_block.cpp:24-99, g_r_k.cpp:28-62, meshes/eval.cpp:62-89,
//   fit_tail_matsubara.cpp:27-72, and test/python/base/tail_issues.py:74-102 (fit_hermitian_tail).
// Needs a GPU; built by __graft_entry__.build() and run by tests/test_gpu_cpp_api.py.
#include <triqs_b200/gfs.hpp>

#include <cstdio>
#include <cstdlib>

using namespace triqs_b200;

static int failures = 0;
#define EXPECT(cond)                                                                                                             \
  do {                                                                                                                           \
    if (!(cond)) {                                                                                                               \
      std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);                                                              \
      ++failures;                                                                                                                \
    }                                                                                                                            \
  } while (0)

template <typename G> double max_diff(G const &a, G const &b) {
  double m = 0;
  for (size_t i = 0; i < a.data().size(); ++i) m = std::max(m, std::abs(a.data()[i] - b.data()[i]));
  return m;
}

static double one_pole(double En, double t, double beta, double s) {
  if (En > 0) return -std::exp(-En * t) / (1 - s * std::exp(-En * beta));
  return s * std::exp(En * (beta - t)) / (1 - s * std::exp(En * beta));
}

template <typename T> void test_fourier(statistic_enum statistic, std::vector<long> shape) {
  double precision = 1e-8, beta = 10, E = -1;
  long N_iw = 1000, N_tau = 10000;
  auto Gw1 = gf<mesh::imfreq, T>{{beta, statistic, N_iw}, shape};
  long no = Gw1.n_others();
  for (long i = 0; i < Gw1.mesh().size(); ++i) {
    dcomplex iw = Gw1.mesh()[i];
    for (long c = 0; c < no; ++c) Gw1[i][c] = 1.0 / (iw - E) + 1.0 / (iw + 2 * E) - 4.5 / (iw - 1.25 * E);
  }
  auto Gt1 = gf<mesh::imtime, T>{{beta, statistic, N_tau}, shape};
  set_from_fourier(Gt1, Gw1);
  auto Gw1b = gf<mesh::imfreq, T>{{beta, statistic, N_iw}, shape};
  set_from_fourier(Gw1b, Gt1);
  EXPECT(max_diff(Gw1, Gw1b) < precision);
  auto Gt1_exact = Gt1;
  double s = (statistic == Fermion ? -1 : 1);
  for (long i = 0; i < Gt1.mesh().size(); ++i) {
    double t = Gt1.mesh()[i];
    for (long c = 0; c < no; ++c) Gt1_exact[i][c] = one_pole(E, t, beta, s) + one_pole(-2 * E, t, beta, s) - 4.5 * one_pole(1.25 * E, t, beta, s);
  }
  EXPECT(max_diff(Gt1, Gt1_exact) < precision);
  // tail from fit_tail passed explicitly
  auto [tail, err] = fit_tail(Gw1);
  auto Gt1_tail = gf<mesh::imtime, T>{{beta, statistic, N_tau}, shape};
  set_from_fourier(Gt1_tail, Gw1, tail);
  EXPECT(max_diff(Gt1_tail, Gt1_exact) < precision);
  // only 0th and 1st moment
  auto known_moments = make_zero_tail(Gt1, 2);
  for (long c = 0; c < no; ++c) known_moments(1, c) = -2.5;
  auto Gt1_km = gf<mesh::imtime, T>{{beta, statistic, N_tau}, shape};
  set_from_fourier(Gt1_km, Gw1, known_moments);
  EXPECT(max_diff(Gt1_km, Gt1_exact) < precision);
  // factories
  auto Gt2 = make_gf_from_fourier(Gw1, N_tau);
  auto Gw2b = make_gf_from_fourier(Gt2, N_iw);
  EXPECT(max_diff(Gt2, Gt1) < precision);
  EXPECT(max_diff(Gw2b, Gw1) < precision);
  // repeated transforms
  for (int i = 0; i < 10; ++i) {
    set_from_fourier(Gt1, Gw1b);
    set_from_fourier(Gw1b, Gt1);
  }
  EXPECT(max_diff(Gw1, Gw1b) < precision);
}

void test_errors() {
  bool thrown = false;
  try {
    auto gt = gf<mesh::imtime, scalar_valued>{{1.0, Fermion, 101}};
    auto gw = make_gf_from_fourier(gt, 60); // time mesh too short
  } catch (runtime_error const &e) { thrown = std::string(e.what()).find("at least twice as long") != std::string::npos; }
  EXPECT(thrown);
}

void test_block() { // fourier_block.cpp: beta = 1, N_iw = 1000, N_tau = 6001
  double beta = 1;
  long N_iw = 1000, N_tau = 6001;
  std::vector<gf<mesh::imfreq, matrix_valued>> blocks;
  for (int b = 0; b < 2; ++b) {
    gf<mesh::imfreq, matrix_valued> g{{beta, Fermion, N_iw}, {2, 2}};
    for (long i = 0; i < g.mesh().size(); ++i) {
      dcomplex iw = g.mesh()[i];
      g[i][0] = 1.0 / (iw - (1.0 + b));
      g[i][3] = 1.0 / (iw + 0.5);
    }
    blocks.push_back(g);
  }
  block_gf<mesh::imfreq, matrix_valued> bgw{{"up", "down"}, blocks};
  auto bgt = make_gf_from_fourier(bgw, N_tau);
  auto bgw2 = make_gf_from_fourier(bgt, N_iw);
  for (int b = 0; b < 2; ++b) EXPECT(max_diff(bgw[b], bgw2[b]) < 1e-7);
}

void test_hermitian_tail() { // tail_issues.py:74-102
  for (double eps : {0.001, 0.1, 1.0, 10.0}) {
    long Nw = std::max(25L, long(eps * 60));
    gf<mesh::imfreq, matrix_valued> g{{1.0, Fermion, Nw}, {2, 2}};
    dcomplex H[2][2] = {{eps * 1.0, eps * dcomplex(0, 0.1)}, {eps * dcomplex(0, -0.1), eps * 2.0}};
    for (long i = 0; i < g.mesh().size(); ++i) {
      dcomplex iw = g.mesh()[i];
      dcomplex a = iw - H[0][0], b = -H[0][1], c = -H[1][0], d = iw - H[1][1];
      dcomplex det = a * d - b * c;
      g[i][0] = d / det; g[i][1] = -b / det; g[i][2] = -c / det; g[i][3] = a / det;
    }
    auto [tail, err] = fit_hermitian_tail(g);
    // moments 1..3 = H^0, H^1, H^2
    dcomplex H2[4] = {H[0][0] * H[0][0] + H[0][1] * H[1][0], H[0][0] * H[0][1] + H[0][1] * H[1][1], H[1][0] * H[0][0] + H[1][1] * H[1][0],
                      H[1][0] * H[0][1] + H[1][1] * H[1][1]};
    dcomplex ex[3][4] = {{1, 0, 0, 1}, {H[0][0], H[0][1], H[1][0], H[1][1]}, {H2[0], H2[1], H2[2], H2[3]}};
    for (int n = 0; n < 3; ++n) {
      double num = 0, den = 0;
      for (int c = 0; c < 4; ++c) { num = std::max(num, std::abs(ex[n][c] - tail(1 + n, c))); den = std::max(den, std::abs(ex[n][c])); }
      EXPECT(num / den < 1e-4);
    }
  }
}

void test_lattice() { // g_r_k.cpp:52-58
  long N = 10;
  gf<mesh::brzone, scalar_valued> gk{{N, N, 1}};
  for (long i = 0; i < N; ++i)
    for (long j = 0; j < N; ++j) gk[i * N + j][0] = -2 * (std::cos(2 * M_PI * i / N) + std::cos(2 * M_PI * j / N));
  auto gr = make_gf_from_fourier(gk);
  EXPECT(std::abs(gr[1 * N + 0][0] + 1.0) < 1e-13);
  EXPECT(std::abs(gr[(N - 1) * N + 0][0] + 1.0) < 1e-13);
  EXPECT(std::abs(gr[0 * N + 1][0] + 1.0) < 1e-13);
  EXPECT(std::abs(gr[0 * N + N - 1][0] + 1.0) < 1e-13);
  EXPECT(std::abs(gr[0][0]) < 1e-13);
  auto gk2 = make_gf_from_fourier(gr);
  EXPECT(max_diff(gk, gk2) < 1e-12);
}

void test_eval_and_inverse() { // meshes/eval.cpp:62-89, gf_inverse.cpp
  gf<mesh::imtime, scalar_valued> g{{10.0, Fermion, 5}};
  double v[5] = {1, 10, 100, 1000, 1e4};
  for (int i = 0; i < 5; ++i) g[i][0] = v[i];
  EXPECT(std::abs(g(2.0)[0] - (0.2 * 1 + 0.8 * 10)) < 1e-13);
  EXPECT(std::abs(g(3.5)[0] - (0.6 * 10 + 0.4 * 100)) < 1e-12);
  gf<mesh::imfreq, matrix_valued> m{{10.0, Fermion, 20}, {1, 1}};
  for (long i = 0; i < m.mesh().size(); ++i) m[i][0] = m.mesh()[i] - 0.3;
  auto mi = inverse(m);
  for (long i = 0; i < m.mesh().size(); ++i) EXPECT(std::abs(mi[i][0] - 1.0 / (m.mesh()[i] - 0.3)) < 1e-14);
  // gf_evaluator<imfreq> (evaluator.hpp:47-75): inside the window the mesh value, outside the fitted tail ~ 1 / (i w - 0.3)
  gf<mesh::imfreq, scalar_valued> gw{{10.0, Fermion, 400}};
  for (long i = 0; i < gw.mesh().size(); ++i) gw[i][0] = 1.0 / (gw.mesh()[i] - 0.3);
  auto ev = evaluate(gw, std::vector<long>{-5000, -400, 0, 7, 399, 400, 3000});
  const double pi = 3.14159265358979323846;
  for (size_t k = 0; k < 7; ++k) {
    const long n = std::vector<long>{-5000, -400, 0, 7, 399, 400, 3000}[k];
    const dcomplex iw{0.0, pi * double(2 * n + 1) / 10.0};
    EXPECT(std::abs(ev[k] - 1.0 / (iw - 0.3)) < (n >= -400 && n <= 399 ? 1e-15 : 1e-7));
  }
  set_option("host_lanes", 2);
  set_option("host_lanes", 3);
  // product mesh: test/c++/gfs/multivar/g_k_om.cpp:76-114 -- g(k, w) re-evaluated on the mesh and on a 3 x finer k grid (within 5 %)
  const long n_k = 40;
  gf<mesh::prod<mesh::brzone, mesh::imfreq>, scalar_valued> gkw{{mesh::brzone{n_k, n_k}, mesh::imfreq{5.0, Fermion, 5}}};
  auto const &wm = gkw.mesh().m2;
  auto fk = [&](double kx, double ky, dcomplex w) { return 1.0 / (w + 3.0 + 2.0 * (std::cos(kx) + std::cos(ky))); };
  for (long ix = 0; ix < n_k; ++ix)
    for (long iy = 0; iy < n_k; ++iy)
      for (long iw = 0; iw < wm.size(); ++iw) gkw[(ix * n_k + iy) * wm.size() + iw][0] = fk(2 * pi * ix / n_k, 2 * pi * iy / n_k, wm[iw]);
  std::vector<std::array<double, 3>> ks;
  std::vector<long> ns;
  std::vector<dcomplex> want;
  const long n2 = 3 * n_k;
  for (long a = 0; a < n2; ++a)
    for (long b = 0; b < n2; ++b)
      for (long n = wm.first_index(); n <= wm.last_index(); ++n) {
        ks.push_back({double(a) / 3.0, double(b) / 3.0, 0.0}); // transpose(units_inv) * k: in units of the mesh index (brzone.hpp:368)
        ns.push_back(n);
        want.push_back(fk(2 * pi * a / n2, 2 * pi * b / n2, wm[n - wm.first_index()]));
      }
  auto got = evaluate(gkw, ks, ns);
  for (size_t i = 0; i < want.size(); ++i) {
    const bool on_mesh = (i / wm.size() / n2) % 3 == 0 && (i / wm.size() % n2) % 3 == 0;
    EXPECT(std::abs(got[i] - want[i]) <= (on_mesh ? 1e-14 : 0.05 * std::abs(want[i])));
  }
}

void test_density() { // gf_density.cpp:22-55, gf_base.cpp:69
  gf<mesh::imfreq, scalar_valued> g{{1.0, Fermion, 400}};
  for (long i = 0; i < g.mesh().size(); ++i) g[i][0] = 1.0 / (g.mesh()[i] + 2.3);
  EXPECT(std::abs(density(g) - 1.0 / (1.0 + std::exp(-2.3))) < 1e-9);
  gf<mesh::imfreq, scalar_valued> gb{{1.0, Boson, 400}};
  for (long i = 0; i < gb.mesh().size(); ++i) gb[i][0] = 1.0 / (gb.mesh()[i] - 2.3);
  EXPECT(std::abs(density(gb) - 1.0 / (-1.0 + std::exp(2.3))) < 1e-9);
  gf<mesh::imfreq, matrix_valued> g3{{1.0, Fermion, 100}, {2, 2}};
  for (long i = 0; i < g3.mesh().size(); ++i) {
    g3[i][0] = g3[i][3] = 2.0 / (g3.mesh()[i] + 2.3);
    g3[i][1] = g3[i][2] = 0.0;
  }
  auto n = density(g3);
  EXPECT(std::abs(n[0] - 1.8177540779781256) < 1e-7 && std::abs(n[3] - 1.8177540779781256) < 1e-7 && std::abs(n[1]) < 1e-12);
}

template <typename T> void test_fourier_real(std::vector<long> shape) { // fourier_real.cpp:27-87
  double precision = 1e-4, w_max = 40, E = 1;
  long Nw = 1001;
  auto w_mesh = mesh::refreq{-w_max, w_max, Nw};
  auto t_mesh = mesh::make_adjoint_mesh(w_mesh);
  auto Gw1 = gf<mesh::refreq, T>{w_mesh, shape};
  long no = Gw1.n_others();
  for (long i = 0; i < Nw; ++i)
    for (long c = 0; c < no; ++c) Gw1[i][c] = 2 * E / (w_mesh[i] * w_mesh[i] + E * E);
  auto tail = make_zero_tail(Gw1, 3);
  for (long c = 0; c < no; ++c) tail(2, c) = 2 * E; // exact expansion of the Lorentzian (the reference fits it)
  auto Gt1 = make_gf_from_fourier(Gw1, t_mesh, tail);
  auto Gw1b = make_gf_from_fourier(Gt1, w_mesh, tail);
  EXPECT(max_diff(Gw1, Gw1b) < precision);
  auto Gt1_exact = Gt1;
  for (long i = 0; i < Nw; ++i)
    for (long c = 0; c < no; ++c) Gt1_exact[i][c] = std::exp(-E * std::abs(t_mesh[i]));
  EXPECT(max_diff(Gt1, Gt1_exact) < precision);
  auto theta = [](double x) { return x > 0 ? 1.0 : (x < 0 ? 0.0 : 0.5); };
  for (long i = 0; i < Nw; ++i) {
    double t = t_mesh[i];
    for (long c = 0; c < no; ++c) Gt1[i][c] = dcomplex(0, 0.5) * (std::exp(-E * std::abs(t)) * theta(-t) - std::exp(-E * std::abs(t)) * theta(t));
  }
  auto known_moments = make_zero_tail(Gt1, 3);
  for (long c = 0; c < no; ++c) known_moments(1, c) = 1.;
  set_from_fourier(Gw1, Gt1, known_moments);
  auto Gt1b = make_gf_from_fourier(Gw1, t_mesh, known_moments);
  EXPECT(max_diff(Gt1, Gt1b) < precision);
  auto Gw1_exact = Gw1;
  for (long i = 0; i < Nw; ++i)
    for (long c = 0; c < no; ++c) Gw1_exact[i][c] = 0.5 / (w_mesh[i] + dcomplex(0, E)) + 0.5 / (w_mesh[i] - dcomplex(0, E));
  EXPECT(max_diff(Gw1, Gw1_exact) < precision);
  // incompatible meshes throw like the reference (fourier_real.cpp:50-51)
  bool thrown = false;
  try {
    make_gf_from_fourier(Gw1, mesh::retime{t_mesh.t_min(), 1.01 * t_mesh.t_max(), Nw});
  } catch (runtime_error const &) { thrown = true; }
  EXPECT(thrown);
}

void test_hermiticity() { // functions/hermiticity.cpp:25-80
  double beta = 1;
  gf<mesh::imfreq, matrix_valued> G{{beta, Fermion, 100}, {2, 2}};
  double H[2][2] = {{1, 0.2}, {0.2, -0.5}};
  for (long i = 0; i < G.mesh().size(); ++i) { // (iw - H)^-1 by hand
    dcomplex a = G.mesh()[i] - H[0][0], b = -H[0][1], c = -H[1][0], d = G.mesh()[i] - H[1][1], det = a * d - b * c;
    G[i][0] = d / det, G[i][1] = -b / det, G[i][2] = -c / det, G[i][3] = a / det;
  }
  EXPECT(is_gf_hermitian(G));
  G[0][1] += dcomplex(0, 1e-3);
  EXPECT(!is_gf_hermitian(G));
  auto Gh = make_hermitian(G);
  EXPECT(is_gf_hermitian(Gh));
  // replace_by_tail: beyond n_min the data becomes the tail expansion
  moment_array tail(4, 4);
  tail(1, 0) = tail(1, 3) = 1.0;
  replace_by_tail(Gh, tail, 50);
  dcomplex iw = Gh.mesh()[Gh.mesh().to_data_index(70)];
  EXPECT(std::abs(Gh[Gh.mesh().to_data_index(70)][0] - 1.0 / iw) < 1e-14);
  EXPECT(std::abs(Gh[Gh.mesh().to_data_index(70)][1]) < 1e-14);
}

void test_legendre_and_tight_binding() { // gf_legendre.cpp:22-72;  g_r_k.cpp:40-58 dispersion
  double beta = 2.0;
  gf<mesh::legendre, scalar_valued> gl{{beta, Fermion, 10}};
  gl[2][0] = 1;
  gf<mesh::imtime, scalar_valued> gt{{beta, Fermion, 5}};
  legendre_matsubara_direct(gt, gl);
  for (long i = 0; i < 5; ++i) {
    double x = 2 * gt.mesh()[i] / beta - 1;
    EXPECT(std::abs(gt[i][0] - std::sqrt(5.0) / beta * (3 * x * x - 1) / 2) < 1e-15);
  }
  gf<mesh::imtime, scalar_valued> gt2{{beta, Fermion, 20001}};
  legendre_matsubara_direct(gt2, gl);
  gf<mesh::legendre, scalar_valued> gl2{{beta, Fermion, 10}};
  legendre_matsubara_inverse(gl2, gt2);
  EXPECT(max_diff(gl, gl2) < 1e-6);
  tight_binding tb;
  tb.n_orb = 1;
  tb.displ_vec = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}};
  tb.overlap_mat_vec = {{-1.0}, {-1.0}, {-1.0}, {-1.0}};
  long N = 10;
  auto ek = tb.fourier(mesh::brzone{N, N, 1});
  for (long i = 0; i < N; ++i)
    for (long j = 0; j < N; ++j) EXPECT(std::abs(ek[i * N + j][0] + 2 * (std::cos(2 * M_PI * i / N) + std::cos(2 * M_PI * j / N))) < 1e-14);
}

// real-valued G(tau): gf<imtime, matrix_real_valued> -> gf<imfreq, matrix_valued> (fourier.hpp:66, :78-81), and real(g) / is_gf_real
// (functions2.hpp:223-246) on the way back
static void test_real_valued_gtau() {
  double beta = 10, E = -1;
  int N_iw = 100, N_tau = 6 * N_iw + 1;
  auto Gt = gf<mesh::imtime, matrix_real_valued>{{beta, Fermion, N_tau}, {2, 2}};
  auto Gtc = gf<mesh::imtime, matrix_valued>{{beta, Fermion, N_tau}, {2, 2}};
  for (long i = 0; i < Gt.mesh().size(); ++i) {
    double t = Gt.mesh()[i];
    for (long c = 0; c < 4; ++c) {
      Gt[i][c] = (one_pole(E, t, beta, -1) + 0.5 * one_pole(-2 * E, t, beta, -1)) * (1 + 0.1 * c);
      Gtc[i][c] = Gt[i][c];
    }
  }
  gf<mesh::imfreq, matrix_valued> Gw = make_gf_from_fourier(Gt, N_iw);
  auto Gwc = make_gf_from_fourier(Gtc, N_iw);
  EXPECT(max_diff(Gw, Gwc) == 0.0); // same bits as promoting on the host
  EXPECT(!is_gf_real(Gw));
  auto Gt_back = make_gf_from_fourier(Gw, N_tau);
  EXPECT(is_gf_real(Gt_back, 1e-6)); // (the fitted tail leaves ~1e-8)
  gf<mesh::imtime, matrix_real_valued> Gr = real(Gt_back);
  auto Gr2 = make_real_gf_from_fourier(Gw, N_tau, {}, 1e-6);
  double d = 0, d2 = 0;
  for (size_t i = 0; i < Gr.data().size(); ++i) {
    d = std::max(d, std::abs(Gr.data()[i] - Gt.data()[i]));
    d2 = std::max(d2, std::abs(Gr.data()[i] - Gr2.data()[i]));
  }
  EXPECT(d < 1e-5); // (round trip through a least-squares tail fit on a short mesh)
  EXPECT(d2 == 0.0);
}

int main() {
  {
    std::string d = plan_describe(mesh::imtime{10.0, Fermion, 5000}, 833, 4, 2); // L = 4999 is prime
    EXPECT(d.find("\"engine_fwd\": \"bluestein\"") != std::string::npos);
  }
  test_real_valued_gtau();
  test_fourier_real<scalar_valued>({});
  test_fourier_real<matrix_valued>({2, 2});
  test_fourier_real<tensor_valued<3>>({2, 2, 2});
  test_hermiticity();
  test_legendre_and_tight_binding();
  test_density();
  test_fourier<scalar_valued>(Fermion, {});
  test_fourier<matrix_valued>(Fermion, {2, 2});
  test_fourier<tensor_valued<3>>(Fermion, {2, 2, 2});
  test_fourier<scalar_valued>(Boson, {});
  test_fourier<matrix_valued>(Boson, {2, 2});
  test_errors();
  test_block();
  test_hermitian_tail();
  test_lattice();
  test_eval_and_inverse();
  std::printf(failures ? "%d FAILURES\n" : "ALL C++ API TESTS PASSED\n", failures);
  return failures ? 1 : 0;
}
