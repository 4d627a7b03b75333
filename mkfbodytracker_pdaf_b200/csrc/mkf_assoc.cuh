// placeholder, replaced below
extern "C" int mkf_batch_associate(mkf_batch*, mkf_batch*, int, const double*, const uint8_t*, const double*, const double*, const double*, const double*, const uint64_t*, int, int) { mkf_set_error("not built yet"); return MKF_E_UNSUPPORTED; }
extern "C" int mkf_batch_assoc_results(mkf_batch*, uint8_t*, double*, int32_t*, int) { mkf_set_error("not built yet"); return MKF_E_UNSUPPORTED; }
