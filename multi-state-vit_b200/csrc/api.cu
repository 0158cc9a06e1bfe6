// Version and error strings of the C ABI (include/msvit.h).
#include "common.cuh"

extern "C" int msvit_version(void) { return 100; }

extern "C" const char* msvit_error_string(int code) {
  switch (code) {
    case MSVIT_OK: return "ok";
    case MSVIT_ERR_NULL: return "a required pointer is NULL";
    case MSVIT_ERR_SHAPE: return "a size is out of the supported range";
    case MSVIT_ERR_ALIGN: return "pointer or row stride is not 16-byte aligned";
    case MSVIT_ERR_MODE: return "unknown dtype or distance mode";
    case MSVIT_ERR_WORKSPACE: return "workspace too small";
    case MSVIT_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown msvit error";
}
