// Integer issue-rate peaks of the B200 for the per-stage rooflines of SURVEY.md §8(d): plain INT32 add/logic, the DPX packed
// u16x2 add-min used by the SGBM path kernels, and POPC (Hamming matcher).  ILP-saturating register-only loops.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak tools/int_peak.cu && ./int_peak > profiles/int_peaks.json
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 4096;

template <int OP>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned seed) {
    unsigned v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = seed + threadIdx.x * 17 + i * 3;
    const unsigned a = seed | 1, b = seed * 3 + 7;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) v[i] = (v[i] + a) ^ b;                        // IADD3 / LOP3 (2 int ops)
            else if (OP == 1) v[i] = __viaddmin_u16x2(v[i], a, b + i); // VIADDMNMX.U16x2 (4 int16 ops: 2 adds + 2 mins)
            else if (OP == 2) v[i] = __popc(v[i] ^ a) + b;             // LOP3 + POPC + IADD
            else if (OP == 3) v[i] = __vminu2(v[i] + a, b);            // IADD + VIMNMX.U16x2
            else if (OP == 4) v[i] = __vminu2(__vmaxu2(v[i], a + i), b + it);  // 2 x VIMNMX.U16x2 (packed min / max alone)
            else v[i] = __vimin3_u16x2(v[i] ^ a, b + i, a + it);       // LOP3 + VIMNMX3.U16x2
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
double run(unsigned* out, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(out, 12345u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k<OP><<<blocks, 256>>>(out, 12345u + r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return (double)blocks * 256 * ITERS * ILP / (best * 1e-3);  // loop bodies per second (per thread-op group)
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8;
    unsigned* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    const double r0 = run<0>(out, blocks), r1 = run<1>(out, blocks), r2 = run<2>(out, blocks), r3 = run<3>(out, blocks);
    const double r4 = run<4>(out, blocks), r5 = run<5>(out, blocks);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"how\": \"tools/int_peak.cu: ILP-%d register-only loops, 8 CTAs x 256 threads per SM, best of 5, CUDA events\",\n",
           p.name, p.multiProcessorCount, ILP);
    printf(" \"int32_add_logic_gops\": %.1f,\n", 2 * r0 / 1e9);
    printf(" \"dpx_viaddmnmx_u16x2_inst_ginst\": %.1f, \"dpx_viaddmnmx_u16x2_int16_gops\": %.1f,\n", r1 / 1e9, 4 * r1 / 1e9);
    printf(" \"popc_xor_add_gops\": %.1f, \"popc_ginst\": %.1f,\n", 3 * r2 / 1e9, r2 / 1e9);
    printf(" \"iadd_vimnmx_u16x2_ginst\": %.1f,\n", 2 * r3 / 1e9);
    printf(" \"vimnmx_u16x2_ginst\": %.1f, \"lop3_vimnmx3_u16x2_ginst\": %.1f}\n", 2 * r4 / 1e9, 2 * r5 / 1e9);
    return 0;
}
