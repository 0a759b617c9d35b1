// placeholder until the device DBPostProcess lands
#include "stages.h"
#include <stdexcept>
namespace b200ocr {
struct DbWork { int unused; };
void db_free(DbWork* w) { delete w; }
void Stages::db_postprocess_device(const float*, int, int, int, const int*, const int*, const b200ocr_det_params&, int,
                                   std::vector<int32_t>*, std::vector<double>*, std::vector<int>*) {
    throw std::runtime_error("unsupported: device DBPostProcess not built yet");
}
}
