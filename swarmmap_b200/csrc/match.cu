// match.cu -- TEMPORARY stubs (replaced by the Hamming matchers); every entry point fails loudly.
#include "swm_internal.cuh"
extern "C" {
int swm_hamming_matrix_device(const uint8_t*, int, const uint8_t*, int, uint16_t*, void*) { return SWM_E_STATE; }
int swm_hamming_matrix(const uint8_t*, int, const uint8_t*, int, uint16_t*, int) { return SWM_E_STATE; }
int swm_hamming_pairs(const uint8_t*, const uint8_t*, int, int32_t*, int) { return SWM_E_STATE; }
int swm_matcher_create(int, swm_matcher**) { return SWM_E_STATE; }
void swm_matcher_destroy(swm_matcher*) {}
const char* swm_matcher_last_error(const swm_matcher*) { return "matchers not built yet"; }
int swm_grid_build(swm_matcher*, const swm_frame_view*, int32_t*, int32_t*) { return SWM_E_STATE; }
int swm_match_init(swm_matcher*, const swm_frame_view*, const swm_frame_view*, float*, int32_t*, int, float, int, int*) { return SWM_E_STATE; }
int swm_match_window(swm_matcher*, const swm_frame_view*, const swm_window_query*, const uint8_t*, int, int, float, int, int32_t*, int*) { return SWM_E_STATE; }
int swm_match_bow(swm_matcher*, const swm_frame_view*, const swm_featvec*, const uint8_t*, const swm_frame_view*, const swm_featvec*, const uint8_t*, int, float, int, int32_t*, int*) { return SWM_E_STATE; }
int swm_db_create(int, const uint8_t*, int64_t, int32_t, int64_t, swm_db**) { return SWM_E_STATE; }
int swm_db_create_device(int, const uint8_t*, int64_t, int32_t, int64_t, swm_db**) { return SWM_E_STATE; }
void swm_db_destroy(swm_db*) {}
int swm_db_query_device(swm_db*, const uint8_t*, int, int, uint64_t*, int32_t*, int, void*) { return SWM_E_STATE; }
int64_t swm_db_size(const swm_db*) { return 0; }
}
