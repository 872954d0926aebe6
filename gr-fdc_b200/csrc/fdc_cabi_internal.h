/* fdc_cabi_internal.h -- shared plumbing of the C-ABI translation units. */
#ifndef FDC_CABI_INTERNAL_H
#define FDC_CABI_INTERNAL_H
#include <cuda_runtime.h>
#include <complex>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/fdc_cabi.h"
#include "fdc_launch.h"
#include "fdc_host.h"

namespace fdc {

void set_error(const std::string& s);
int fail(const std::string& s);                       /* records the text, returns -1 */
int cuda_fail(cudaError_t e, const char* what);       /* same with the CUDA error string */
bool require_device();                                /* false (+ error text) when no CUDA device is usable */

/* A context lives on the device that was current when it was made.  CUDA's current device is a per-thread setting and GNU
 * Radio calls a block's work() from a scheduler thread, not from the thread that constructed the block: every entry point
 * that touches the device switches to the context's device for the duration of the call. */
struct OnDevice {
    int prev;
    explicit OnDevice(int dev) : prev(-1)
    {
        int cur = -1;
        if (dev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != dev && cudaSetDevice(dev) == cudaSuccess) prev = cur;
    }
    ~OnDevice() { if (prev >= 0) cudaSetDevice(prev); }
private:
    OnDevice(const OnDevice&); OnDevice& operator=(const OnDevice&);
};
struct DevCtx { int dev; DevCtx() : dev(-1) { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } } };

/* grow-only device buffer */
struct DevBuf {
    void* p; size_t cap;
    DevBuf() : p(0), cap(0) {}
    ~DevBuf() { if (p) cudaFree(p); }
    bool reserve(size_t bytes)
    {
        if (bytes == 0) bytes = 16;
        if (bytes <= cap) return true;
        if (p) { cudaFree(p); p = 0; cap = 0; }
        if (cudaMalloc(&p, bytes) != cudaSuccess) { p = 0; return false; }
        cap = bytes; return true;
    }
    bool upload(const void* src, size_t bytes)
    {
        if (!reserve(bytes)) return false;
        if (bytes == 0) return true;
        return cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    void swap(DevBuf& o) { void* tp = p; p = o.p; o.p = tp; size_t tc = cap; cap = o.cap; o.cap = tc; }
private:
    DevBuf(const DevBuf&); DevBuf& operator=(const DevBuf&);
};

/* grow-only pinned host buffer */
struct PinBuf {
    void* p; size_t cap;
    PinBuf() : p(0), cap(0) {}
    ~PinBuf() { if (p) cudaFreeHost(p); }
    bool reserve(size_t bytes)
    {
        if (bytes == 0) bytes = 16;
        if (bytes <= cap) return true;
        if (p) { cudaFreeHost(p); p = 0; cap = 0; }
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { p = 0; return false; }
        cap = bytes; return true;
    }
private:
    PinBuf(const PinBuf&); PinBuf& operator=(const PinBuf&);
};

}  // namespace fdc
#endif
