// batch.cu -- batch entry point (include/qpalm_b200.h Part 3).
#include "../../include/qpalm_b200.h"
#include "engine.cuh"
extern "C" QPALMB200Batch *qpalm_b200_batch_setup(const QPALMData *, const QPALMSettings *, c_int) { return nullptr; }
extern "C" int qpalm_b200_batch_solve(QPALMB200Batch *, c_int, const c_float *, const c_float *, const c_float *, c_float *, c_float *, QPALMInfo *) { return 1; }
extern "C" int qpalm_b200_batch_upload(QPALMB200Batch *, c_int, const c_float *, const c_float *, const c_float *) { return 1; }
extern "C" int qpalm_b200_batch_solve_resident(QPALMB200Batch *, c_int, double *) { return 1; }
extern "C" int qpalm_b200_batch_download(QPALMB200Batch *, c_int, c_float *, c_float *, QPALMInfo *) { return 1; }
extern "C" void qpalm_b200_batch_cleanup(QPALMB200Batch *) {}
