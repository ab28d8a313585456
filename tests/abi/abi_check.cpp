// abi_check.cpp -- TEST: a plain C++ translation unit compiled ONLY against include/iamrx.h and linked with
// libiamrx.so (no ctypes, no torch): what an IAMR maintainer's adapter would see.  Exit code 0 = ok.
//   abi_check host   : host-side entry points + the loud no-device error of the compute entries (CPU container)
//   abi_check device : one Taylor-Green step through iamrx_ns_* on the GPU and two discrete invariants
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "iamrx.h"

#define REQUIRE(c) do { if (!(c)) { std::fprintf(stderr, "abi_check: %s failed (line %d): %s\n", #c, __LINE__, iamrx_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  const bool device = argc > 1 && std::strcmp(argv[1], "device") == 0;
  static_assert(sizeof(iamrx_box) == 24, "iamrx_box layout");
  static_assert(sizeof(iamrx_fab) == 64, "iamrx_fab layout");
  static_assert(sizeof(iamrx_bcrec) == 24, "iamrx_bcrec layout");
  REQUIRE(iamrx_version() >= 100);
  const int n = 16;
  iamrx_geom g{};
  for (int d = 0; d < 3; ++d) { g.domain.lo[d] = 0; g.domain.hi[d] = n - 1; g.dx[d] = 1.0 / n; g.prob_lo[d] = 0.0; g.periodic[d] = 1; }
  iamrx_box boxes[2] = {{{0, 0, 0}, {n / 2 - 1, n - 1, n - 1}}, {{n / 2, 0, 0}, {n - 1, n - 1, n - 1}}};
  int owner[2] = {0, 0};
  iamrx_level_t lev = nullptr;
  REQUIRE(iamrx_level_create(&g, 2, boxes, owner, &lev) == IAMRX_OK);
  REQUIRE(iamrx_level_num_local(lev) == 2);
  iamrx_box b{}; int gi = -1;
  REQUIRE(iamrx_level_local_box(lev, 1, &b, &gi) == IAMRX_OK && gi == 1 && b.lo[0] == n / 2);
  REQUIRE(iamrx_debug_fb_plan(lev, 0, 1, 0, nullptr, nullptr, nullptr, nullptr) > 0);   // pure host logic
  // overlapping boxes are rejected with a message, nothing throws across the boundary
  iamrx_box bad[2] = {boxes[0], boxes[0]};
  iamrx_level_t lev2 = nullptr;
  REQUIRE(iamrx_level_create(&g, 2, bad, owner, &lev2) == IAMRX_ERR_ARG && std::strstr(iamrx_last_error(), "overlap"));
  iamrx_ns_params p; iamrx_ns_params_default(&p);
  REQUIRE(p.cfl == 0.7 && p.be_cn_theta == 0.5 && p.init_iter == 2);
  p.visc_coef = 1e-3;
  iamrx_mg_info mi; iamrx_mg_info_default(&mi);
  REQUIRE(mi.nu1 == 2 && mi.nu2 == 2 && mi.max_iter == 200);
  iamrx_ns_t ns = nullptr;
  const int rc = iamrx_ns_create(lev, &p, &ns);
  if (!device) {
    // no CUDA device: every compute entry fails loudly, there is no CPU path
    if (iamrx_device_ok()) { std::printf("abi_check host: a device is present, skipping the no-device assertions\n"); }
    else {
      REQUIRE(rc == IAMRX_ERR_NO_DEVICE && std::strstr(iamrx_last_error(), "no CPU fallback"));
      double x[3] = {1, 1, 1};
      REQUIRE(iamrx_abec_gsrb_box(&boxes[0], nullptr, nullptr, 0, 1, nullptr, nullptr, nullptr, nullptr, x, 1.0, 0, 1, nullptr) == IAMRX_ERR_NO_DEVICE);
      REQUIRE(iamrx_compute_aofs_box(&boxes[0], nullptr, 0, nullptr, 0, 1, nullptr, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                     nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, &g, 0.1, 0, nullptr) == IAMRX_ERR_NO_DEVICE);
    }
    iamrx_level_destroy(lev);
    std::printf("abi_check host ok\n");
    return 0;
  }
  REQUIRE(rc == IAMRX_OK);
  const double tg[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
  REQUIRE(iamrx_ns_init_prob(ns, 11, tg, 5) == IAMRX_OK);
  double dt0 = 0.0;
  REQUIRE(iamrx_ns_post_init(ns, &dt0) == IAMRX_OK && dt0 > 0.0);
  double q0[3], q1[3];
  REQUIRE(iamrx_ns_sum_integrated_quantities(ns, q0) == IAMRX_OK);
  double dt = -1.0;
  REQUIRE(iamrx_ns_step(ns, &dt) == IAMRX_OK && std::fabs(dt - dt0) <= 1e-14);
  REQUIRE(iamrx_ns_sum_integrated_quantities(ns, q1) == IAMRX_OK);
  REQUIRE(std::fabs(q1[0] - q0[0]) <= 1e-12);        // mass is conserved
  REQUIRE(q1[2] <= q0[2] * (1.0 + 1e-12));            // viscous flow: kinetic energy does not grow
  int it[3];
  REQUIRE(iamrx_ns_last_iters(ns, it) == IAMRX_OK && it[0] > 0 && it[2] > 0);
  REQUIRE(iamrx_launch_count() > 0);
  iamrx_ns_destroy(ns);
  iamrx_level_destroy(lev);
  std::printf("abi_check device ok: dt %.6e, MG iterations mac/visc/nodal %d/%d/%d\n", dt, it[0], it[1], it[2]);
  return 0;
}
