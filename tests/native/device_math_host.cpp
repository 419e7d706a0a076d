// Host build of h263_rs_b200/csrc/device_math.cuh for the CPU self-test
// (tests/test_device_math.py).  Compiled with g++ -ffp-contract=off; exposes the HD
// primitives through a C ABI so they can be compared with the oracle without a GPU.
#include "../../h263_rs_b200/csrc/device_math.cuh"

using namespace h263dev;

static const float BASIS[8][8] = H263_BASIS_TABLE;
static const unsigned char DEZIGZAG[64] = H263_DEZIGZAG_LINEAR;
static const unsigned char Q2S[32] = H263_QUANT_TO_STRENGTH;

extern "C" {
int dm_dequant(int level, int quant) { return dequant(level, quant); }
int dm_intradc_level(int code) { return intradc_level(code); }
int dm_round_residual(float v) { return round_residual(v); }
int dm_round_residual_scaled(float v) { return round_residual_scaled(v); }
int dm_round_residual_dc(float v) { return round_residual_dc(v); }
unsigned dm_avg2(unsigned a, unsigned b) { return avg2_u8x4(a, b); }
unsigned dm_avg4(unsigned a, unsigned b, unsigned c, unsigned d) { return avg4_u8x4(a, b, c, d); }
int dm_average_sum_of_mvs(int s) { return average_sum_of_mvs(s); }
unsigned dm_yuv_pixel(int y, int cb, int cr) { return yuv_pixel(y, chroma_terms(cb, cr)); }
void dm_deblock_process(int* abcd, int strength, int trunc) {
    deblock_process(abcd[0], abcd[1], abcd[2], abcd[3], strength, trunc);
}
float dm_basis(int f, int i) { return BASIS[f][i]; }
int dm_dezigzag(int p) { return DEZIGZAG[p]; }
int dm_q2s(int q) { return Q2S[q]; }

// The kernel's block transform, lane by lane, in the kernel's order of operations:
// rows mask -> row pass over non-empty rows -> column pass over the same rows -> rounding.
// cls: 3 = Vert, 4 = Full (Horiz is computed as Full).  coefs row-major [y*8+x].
void dm_block_transform(int cls, const float* coefs, unsigned rows, int* residual /*[y*8+x]*/) {
    float t[8][8];
    for (int y = 0; y < 8; y++)
        for (int i = 0; i < 8; i++) t[y][i] = 0.0f;
    for (int y = 0; y < 8; y++) {
        if (!((rows >> y) & 1u)) continue;
        for (int i = 0; i < 8; i++) {
            if (cls == 3) {
                t[y][i] = coefs[y * 8];
            } else {
                float acc = 0.0f;
                for (int x = 0; x < 8; x++) acc = fadd(acc, fmul(coefs[y * 8 + x], BASIS[x][i]));
                t[y][i] = acc;
            }
        }
    }
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            float acc = 0.0f;
            for (int y = 0; y < 8; y++)
                if ((rows >> y) & 1u) acc = fadd(acc, fmul(t[y][i], BASIS[y][j]));
            residual[j * 8 + i] = cls == 3 ? round_residual_scaled(acc) : round_residual(acc);
        }
}
}
