// Handle management and the small C-ABI entry points of libfbdsp.so.
#include "common.cuh"
#include "psk_shared.cuh"

int fb_ensure(fb_handle* h, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return FB_OK;
  // grow geometrically so a stream of slightly larger batches does not reallocate every call
  size_t want = bytes + bytes / 4 + 256;
  if (b.p) {
    FB_CUDA(h, cudaStreamSynchronize(h->stream));
    FB_CUDA(h, cudaStreamSynchronize(h->stream2));
    FB_CUDA(h, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&b.p, want);
  }
  if (e != cudaSuccess) {
    h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    b.p = nullptr;
    return FB_ENOMEM;
  }
  b.cap = want;
  return FB_OK;
}

extern "C" int fb_abi_version(void) { return FB_ABI_VERSION; }

extern "C" int fb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" const char* fb_strerror(int code) {
  switch (code) {
    case FB_OK: return "ok";
    case FB_EINVAL: return "invalid argument";
    case FB_ECUDA: return "CUDA failure (no usable sm_100 device, or a runtime error; see fb_last_error)";
    case FB_ENOMEM: return "out of device memory";
    case FB_EUNSUPPORTED: return "parameter set not supported by the device path";
    default: return "unknown error";
  }
}

extern "C" const char* fb_last_error(fb_handle* h) { return h ? h->err.c_str() : "null handle"; }

extern "C" fb_handle* fb_create(int device) {
  int n = fb_device_count();
  if (device < 0 || device >= n) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return nullptr;
  if (prop.major != 10) {   // the fatbin holds sm_100a code only: fail loudly instead of at the first launch
    fprintf(stderr, "libfbdsp: device %d is sm_%d%d, this library is built for sm_100a only\n", device, prop.major, prop.minor);
    return nullptr;
  }
  fb_handle* h = new fb_handle();
  for (auto& e : h->ev_copy)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { delete h; return nullptr; }
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  // stream2 carries the small latency-bound edge kernels that run beside the big interior kernels: it gets the highest
  // priority so its few CTAs are placed as soon as resources free up instead of after the interior grid has drained
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->stream_copy, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&h->ev_k0) != cudaSuccess || cudaEventCreate(&h->ev_k1) != cudaSuccess) {
    delete h;
    return nullptr;
  }
  return h;
}

extern "C" void fb_destroy(fb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->stream2);
  fb_fsk_release(h);
  fb_resample_release(h);
  fb_psk_mma_release(h);
  DevBuf* bufs[] = {&h->in, &h->out, &h->out_len, &h->sync_idx, &h->status, &h->bits, &h->plans, &h->tile_first, &h->tiles,
                    &h->jobs, &h->scratch, &h->taps, &h->slow_w, &h->sync_raw, &h->fec_in, &h->fec_out, &h->fec_meta, &h->misc, &h->redo, &h->mma_trace, &h->fftws, &h->psk_tabs};
  for (DevBuf* b : bufs)
    if (b->p) cudaFree(b->p);
  cudaEventDestroy(h->ev_fork);
  cudaEventDestroy(h->ev_join);
  cudaEventDestroy(h->ev_k0);
  cudaEventDestroy(h->ev_k1);
  cudaStreamDestroy(h->stream);
  cudaStreamDestroy(h->stream2);
  if (h->stream_copy) cudaStreamDestroy(h->stream_copy);
  for (auto& e : h->ev_copy) if (e) cudaEventDestroy(e);
  delete h;
}

extern "C" void* fb_stream(fb_handle* h) { return h ? (void*)h->stream : nullptr; }

extern "C" int fb_sync(fb_handle* h) {
  if (!h) return FB_EINVAL;
  FB_LOCK(h);
  FB_CUDA(h, cudaSetDevice(h->device));
  FB_CUDA(h, cudaStreamSynchronize(h->stream));
  FB_CUDA(h, cudaStreamSynchronize(h->stream2));
  return FB_OK;
}

extern "C" uint64_t fb_kernel_launches(fb_handle* h) { return h ? h->launches : 0; }

extern "C" int fb_set_profiling(fb_handle* h, int on) {
  if (!h) return FB_EINVAL;
  h->profiling = on != 0;
  h->k_recorded = false;
  return FB_OK;
}

extern "C" float fb_kernel_ms(fb_handle* h) {
  if (!h) return -1.f;
  FB_LOCK(h);
  if (!h->k_recorded) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(h->ev_k1) != cudaSuccess) return -1.f;
  if (cudaEventElapsedTime(&ms, h->ev_k0, h->ev_k1) != cudaSuccess) return -1.f;
  return ms;
}
