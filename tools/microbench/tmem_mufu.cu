// Microbenchmarks behind the attention analysis in DESIGN.md (run on a B200: tools/_trace/tmem_mufu):
//   (1) tcgen05.ld 32x32b.x32 throughput: 1 warp, 4 warps on the 4 lane quarters, 8 warps (2 per quarter)
//   (2) ex2.approx throughput of ONE warp per SM sub-partition (independent streams) vs two warps per sub-partition
//   (3) the softmax inner loop of attn.cu (FFMA2 + 2 x MUFU + FADD2 + F2FP per score pair) on registers only
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_trace/tmem_mufu tools/microbench/tmem_mufu.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x128(uint32_t taddr, uint32_t (&r)[128]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x128.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64, %65, %66, %67, %68, %69, %70, %71, %72, %73, %74, %75, %76, %77, %78, %79, %80, %81, %82, %83, %84, %85, %86, %87, %88, %89, %90, %91, %92, %93, %94, %95, %96, %97, %98, %99, %100, %101, %102, %103, %104, %105, %106, %107, %108, %109, %110, %111, %112, %113, %114, %115, %116, %117, %118, %119, %120, %121, %122, %123, %124, %125, %126, %127}, [%128];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]), "=r"(r[64]), "=r"(r[65]), "=r"(r[66]), "=r"(r[67]), "=r"(r[68]), "=r"(r[69]), "=r"(r[70]), "=r"(r[71]), "=r"(r[72]), "=r"(r[73]), "=r"(r[74]), "=r"(r[75]), "=r"(r[76]), "=r"(r[77]), "=r"(r[78]), "=r"(r[79]), "=r"(r[80]), "=r"(r[81]), "=r"(r[82]), "=r"(r[83]), "=r"(r[84]), "=r"(r[85]), "=r"(r[86]), "=r"(r[87]), "=r"(r[88]), "=r"(r[89]), "=r"(r[90]), "=r"(r[91]), "=r"(r[92]), "=r"(r[93]), "=r"(r[94]), "=r"(r[95]), "=r"(r[96]), "=r"(r[97]), "=r"(r[98]), "=r"(r[99]), "=r"(r[100]), "=r"(r[101]), "=r"(r[102]), "=r"(r[103]), "=r"(r[104]), "=r"(r[105]), "=r"(r[106]), "=r"(r[107]), "=r"(r[108]), "=r"(r[109]), "=r"(r[110]), "=r"(r[111]), "=r"(r[112]), "=r"(r[113]), "=r"(r[114]), "=r"(r[115]), "=r"(r[116]), "=r"(r[117]), "=r"(r[118]), "=r"(r[119]), "=r"(r[120]), "=r"(r[121]), "=r"(r[122]), "=r"(r[123]), "=r"(r[124]), "=r"(r[125]), "=r"(r[126]), "=r"(r[127])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
      "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}

// mode 0: TMEM loads only (4 x32 loads of one 128-column block per iteration, all issued, one wait)
// mode 1: MUFU only: 128 ex2 per thread per iteration from registers
// mode 2: the attention inner loop on registers (no TMEM): per 32 values FFMA2 x16, MUFU x32, FADD2 x16, pack x16
// mode 3: TMEM loads (double-buffered as in attn.cu) + the inner loop
__global__ void __launch_bounds__(256, 1) bench(int mode, int warps_active, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  uint32_t pk = 0;
  const float2 sc = make_float2(1.0001f, 1.0001f), nm = make_float2(-0.5f, -0.5f);
  auto process = [&](const uint32_t (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 x = ffma2(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc, nm);
      const float2 e = make_float2(ex2(x.x), ex2(x.y));
      acc[i & 1] = fadd2(acc[i & 1], e);
      __nv_bfloat162 h = __floats2bfloat162_rn(e.x, e.y);
      pk ^= *reinterpret_cast<uint32_t*>(&h);
    }
  };
  long long t0 = 0, t1 = 0;
  if (warp < warps_active) {
    uint32_t va[32], vb[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { va[i] = __float_as_uint(-1.0f - 0.01f * (i + lane)); vb[i] = __float_as_uint(-2.0f - 0.01f * (i + lane)); }
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
        tmem_ld_x32(base + 0, va); tmem_ld_x32(base + 32, vb);
        tmem_ld_wait();
        acc[0].x += __uint_as_float(va[it & 31]) + __uint_as_float(vb[it & 31]);
        tmem_ld_x32(base + 64, va); tmem_ld_x32(base + 96, vb);
        tmem_ld_wait();
        acc[0].y += __uint_as_float(va[it & 31]) + __uint_as_float(vb[it & 31]);
      } else if (mode == 4) {
        uint32_t w[64];
        tmem_ld_x64(base + 0, w);
        tmem_ld_wait();
        acc[0].x += __uint_as_float(w[it & 63]);
        tmem_ld_x64(base + 64, w);
        tmem_ld_wait();
        acc[0].y += __uint_as_float(w[it & 63]);
      } else if (mode == 5) {
        uint32_t w[128];
        tmem_ld_x128(base + 0, w);
        tmem_ld_wait();
        acc[0].x += __uint_as_float(w[it & 127]);
      } else if (mode == 6) {
        uint32_t w[64];
        tmem_ld_x64(base + 0, w);
        tmem_ld_wait();
        process(*reinterpret_cast<uint32_t(*)[32]>(&w[0]));
        process(*reinterpret_cast<uint32_t(*)[32]>(&w[32]));
        tmem_ld_x64(base + 64, w);
        tmem_ld_wait();
        process(*reinterpret_cast<uint32_t(*)[32]>(&w[0]));
        process(*reinterpret_cast<uint32_t(*)[32]>(&w[32]));
      } else if (mode == 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int i = 0; i < 32; ++i) va[i] = __float_as_uint(ex2(__uint_as_float(va[i])) - 1.5f);
      } else if (mode == 2) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
#pragma unroll
          for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(va[i]), "+r"(vb[i]));   // opaque: no hoisting of loop-invariant exp2
          process(va);
          process(vb);
        }
      } else {
        tmem_ld_x32(base + 0, va);
        tmem_ld_wait();
        tmem_ld_x32(base + 32, vb);
        process(va);
        tmem_ld_wait();
        tmem_ld_x32(base + 64, va);
        process(vb);
        tmem_ld_wait();
        tmem_ld_x32(base + 96, vb);
        process(va);
        tmem_ld_wait();
        process(vb);
      }
    }
    t1 = clock64();
    float s = acc[0].x + acc[0].y + acc[1].x + acc[1].y + __uint_as_float(va[lane]) + __uint_as_float(pk);
    if (s == 123.456f) sink[threadIdx.x] = s;
  }
  if (lane == 0 && blockIdx.x == 0) out[warp] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
  }
}

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 8 * sizeof(long long)); cudaMalloc(&d_sink, 256 * sizeof(float));
  const char* names[7] = {"tcgen05.ld x32 only (128 fp32 columns per iteration)", "MUFU.EX2 only (128 per thread per iteration)",
                          "attention inner loop on registers (128 scores per iteration)", "TMEM loads + inner loop (as in attn.cu)",
                          "tcgen05.ld x64 only (2 loads per block)", "tcgen05.ld x128 only (1 load per block)",
                          "x64 loads (2 per block) + inner loop"};
  const int iters = 2000;
  for (int mode = 0; mode < 7; ++mode)
    for (int w : {1, 4, 8}) {
      cudaMemset(d_out, 0, 8 * sizeof(long long));
      bench<<<1, 256>>>(mode, w, iters, d_out, d_sink);   // warps 0-3: one per sub-partition / lane quarter; 4-7: a second one each
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d warps %d: %s\n", mode, w, cudaGetErrorString(e)); return 1; }
      long long h[8]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < w; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("%-62s %d warp(s): %7.1f cycles per 128-column block per warp\n", names[mode], w, (double)mx / iters);
    }
  return 0;
}
