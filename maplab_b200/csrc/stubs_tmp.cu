#include "detector.h"
namespace mlc {
bool Detector::PnpRansacBatch(const mlc_ransac_settings&, const mlc_camera*, int, int64_t, const int64_t*, const double*, const int32_t*, const int32_t*, const double*, mlc_pose_result*, uint8_t*, std::string* err) { *err = "not built yet"; return false; }
}
