// placeholder, replaced below
struct mkf_pf2d { int dummy; };
extern "C" int mkf_pf2d_create(mkf_pf2d**, int64_t, int, int, int, const double*, const double*, const double*, int, void*) { mkf_set_error("not built yet"); return MKF_E_UNSUPPORTED; }
extern "C" void mkf_pf2d_destroy(mkf_pf2d*) {}
extern "C" int mkf_pf2d_set_particles(mkf_pf2d*, const double*, int) { return MKF_E_UNSUPPORTED; }
extern "C" int mkf_pf2d_get(mkf_pf2d*, double*, double*, int32_t*, int) { return MKF_E_UNSUPPORTED; }
extern "C" int mkf_pf2d_update(mkf_pf2d*, const double*, const double*, const double*, int) { return MKF_E_UNSUPPORTED; }
extern "C" int mkf_pf2d_sync(mkf_pf2d*) { return MKF_E_UNSUPPORTED; }
