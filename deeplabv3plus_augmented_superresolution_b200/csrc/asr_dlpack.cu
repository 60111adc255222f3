// asr_dlpack.cu -- DLPack front door of the C ABI (north_star: "Python host code calls a thin C-ABI
// shared library via ctypes over DLPack buffers").  The structs below restate the public dlpack.h
// v0.8 ABI (DLDevice / DLDataType / DLTensor); a DLManagedTensor* obtained from a "dltensor"
// PyCapsule can be passed directly because DLTensor is its first member.
#include "asr_common.cuh"

extern "C" {
typedef struct { int32_t device_type; int32_t device_id; } AsrDLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } AsrDLDataType;
struct DLTensor {
    void* data;
    AsrDLDevice device;
    int32_t ndim;
    AsrDLDataType dtype;
    int64_t* shape;
    int64_t* strides;   // NULL = compact row-major
    uint64_t byte_offset;
};
}

namespace {
constexpr int kDLCUDA = 2, kDLUInt = 1, kDLFloat = 2;

int validate(const DLTensor* t, const char* name, int code, int bits, int ndim) {
    if (!t) return asr::fail(ASR_ENULL, "%s is NULL", name);
    if (t->device.device_type != kDLCUDA) return asr::fail(ASR_EDTYPE, "%s: device type %d is not kDLCUDA", name, t->device.device_type);
    if (t->dtype.code != code || t->dtype.bits != bits || t->dtype.lanes != 1)
        return asr::fail(ASR_EDTYPE, "%s: dtype (%d,%d,%d) unexpected", name, t->dtype.code, t->dtype.bits, t->dtype.lanes);
    if (t->ndim != ndim) return asr::fail(ASR_EINVAL, "%s: ndim %d, expected %d", name, t->ndim, ndim);
    if (t->strides) {
        int64_t expect = 1;
        for (int i = t->ndim - 1; i >= 0; --i) {
            if (t->shape[i] != 1 && t->strides[i] != expect) return asr::fail(ASR_EDTYPE, "%s: not C-contiguous", name);
            expect *= t->shape[i];
        }
    }
    return ASR_OK;
}
void* ptr(const DLTensor* t) { return static_cast<char*>(t->data) + t->byte_offset; }
}  // namespace

extern "C" int asr_solve_batched_dlpack(const AsrSolveParams* params, int n_params, const DLTensor* copies,
                                        const float* h_angles, const float* h_shifts, const uint8_t* h_keep,
                                        DLTensor* x_out, DLTensor* loss_out, DLTensor* workspace, void* stream) {
    if (int e = validate(copies, "copies", kDLFloat, 32, 4)) return e;
    if (int e = validate(x_out, "x_out", kDLFloat, 32, 3)) return e;
    if (int e = validate(workspace, "workspace", kDLUInt, 8, 1)) return e;
    if (loss_out) if (int e = validate(loss_out, "loss_out", kDLFloat, 32, 1)) return e;
    const int64_t B = copies->shape[0], N = copies->shape[1], h = copies->shape[2], w = copies->shape[3];
    if (x_out->shape[0] != B) return asr::fail(ASR_EINVAL, "x_out batch %lld != copies batch %lld", (long long)x_out->shape[0], (long long)B);
    if (loss_out && loss_out->shape[0] != B) return asr::fail(ASR_EINVAL, "loss_out must have B entries");
    if (x_out->device.device_id != copies->device.device_id || workspace->device.device_id != copies->device.device_id)
        return asr::fail(ASR_EINVAL, "tensors live on different devices");
    return asr_solve_batched(params, n_params, static_cast<const float*>(ptr(copies)), h_angles, h_shifts, h_keep, (int)B,
                             (int)N, (int)h, (int)w, (int)x_out->shape[1], (int)x_out->shape[2],
                             static_cast<float*>(ptr(x_out)), loss_out ? static_cast<float*>(ptr(loss_out)) : nullptr,
                             ptr(workspace), (size_t)workspace->shape[0], stream);
}
