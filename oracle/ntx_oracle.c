/* oracle/ntx_oracle.c — CPU restatement of the reference's per-ray-sample hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (nerf_texture_b200/, include/) may link, load or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do, as the checker / the reported CPU baseline.
 *
 * What it restates (reference = yihua7/NeRF-Texture, paths relative to /root/reference):
 *   gridencoder/src/gridencoder.cu   fast_hash:36  get_grid_index:55  kernel_grid:76
 *                                    kernel_grid_backward:228  kernel_input_backward:318
 *   ffmlp/src/ffmlp.cu               kernel_mlp_fused:332 (+ layouts :631-634)  kernel_mlp_fused_backward:411
 *                                    ffmlp_backward:749 (weight-gradient GEMMs)  utils.h:425,538 (activations)
 *   shencoder/src/shencoder.cu       kernel_sh:28  kernel_sh_backward:360
 *   raymarching/src/raymarching.cu   :44-83 (mip / morton)  :94 near_far  :165 polar  :270 packbits
 *                                    :314 march_rays_train  :700/:802 composite_rays_train fwd/bwd
 *                                    :900 march_rays  :1021 composite_rays  :1117 compact_rays
 *   raymarching/src/pcg32.h          :57-72 seed/next_uint  :107-116 next_float  :149-170 advance
 *
 * The reference has NO CPU implementation of any of these (every native entry point CHECK_CUDAs its
 * tensors) and no golden vectors, so this restatement is pinned against the reference's own CUDA
 * kernels recompiled for sm_100a (oracle/_ref, built by oracle/build_ref.py) on the GPU box, and against
 * the fixtures under tests/golden/ that were produced by those kernels (tests/golden/make_golden.py).
 *
 * Numerics notes (deliberate, documented differences from bit-exactness):
 *   - exp2f / expf / atan2f here are glibc's; the GPU uses ex2.approx-based exp2f and __expf.  The per-level
 *     grid scale can be injected (level_scales) so that integer index streams are compared bit-exactly.
 *   - a*b+c patterns that nvcc contracts to FMA under its default -fmad=true are written as fmaf().
 *   - fp16 table arithmetic rounds where c10::Half's operators round (see grid_impl.inc).
 *   - the fused MLP here accumulates each dot product in fp32 and rounds activations to fp16 once per
 *     layer (acc_mode 0, what the B200 kernel does) or additionally rounds the accumulator to fp16 after
 *     every 16-wide K block (acc_mode 1, an approximation of the reference's wmma fp16 accumulators).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 orc_half;
static inline float h2f(orc_half h) { return (float)h; }
static inline orc_half f2h(float f) { return (orc_half)f; }

/* ------------------------------------------------------------------------------------------------
 *                                         grid encoder
 * ---------------------------------------------------------------------------------------------- */

/* gridencoder.cu:126: scale = exp2f(level * S) * H - 1.0f  (the multiply-subtract contracts to one FMA) */
static inline float orc_level_scale(uint32_t level, float S, uint32_t H)
{
    return fmaf(exp2f((float)level * S), (float)H, -1.0f);
}

/* gridencoder.cu:36-51 */
static inline uint32_t orc_fast_hash(const uint32_t* p, uint32_t D)
{
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
    for (uint32_t i = 0; i < D; i++) r ^= p[i] * primes[i];
    return r;
}

/* gridencoder.cu:55-72 */
static inline uint32_t orc_grid_index(uint32_t gridtype, int align, uint32_t D, uint32_t C, uint32_t ch,
                                      uint32_t hashmap_size, uint32_t resolution, const uint32_t* pg)
{
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pg[d] * stride;
        stride *= align ? resolution : (resolution + 1);
    }
    if (gridtype == 0 && stride > hashmap_size) index = orc_fast_hash(pg, D);
    return (index % hashmap_size) * C + ch;
}

void orc_grid_level_scales(float S, uint32_t H, uint32_t L, float* out)
{
    for (uint32_t l = 0; l < L; l++) out[l] = orc_level_scale(l, S, H);
}

/* integer index stream of one level: idx[b*2^D + corner] — used to diff the GPU's indices bit-exactly */
int orc_grid_indices(const float* inputs, const int* offsets, uint32_t B, uint32_t D, uint32_t level, float scale,
                     uint32_t gridtype, int align, uint32_t* idx_out)
{
    const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    const uint32_t resolution = (uint32_t)ceilf(scale) + 1;
    for (uint32_t b = 0; b < B; b++) {
        const float* x = inputs + (size_t)b * D;
        uint32_t pg[3];
        int oob = 0;
        for (uint32_t d = 0; d < D; d++) if (x[d] < 0 || x[d] > 1) oob = 1;
        for (uint32_t d = 0; d < D; d++) pg[d] = oob ? 0 : (uint32_t)floorf(fmaf(x[d], scale, align ? 0.0f : 0.5f));
        for (uint32_t c = 0; c < (1u << D); c++) {
            uint32_t pl[3];
            for (uint32_t d = 0; d < D; d++) pl[d] = pg[d] + ((c >> d) & 1u);
            idx_out[(size_t)b * (1u << D) + c] = oob ? 0xffffffffu : orc_grid_index(gridtype, align, D, 1, 0, hashmap_size, resolution, pl);
        }
    }
    return 0;
}

/* fp32 tables */
#define SCALAR float
#define FN(n) n##_f32
#define S_ZERO 0.0f
#define S_ACC(a, w, v) ((a) = fmaf((w), (v), (a)))
#define S_SUB(r, l) ((r) - (l))
#define S_WMUL(w, v) ((w) * (v))
#define S_ADD(a, b) ((a) + (b))
#define S_MULACC(a, x, y) ((a) = fmaf((x), (y), (a)))
#include "grid_impl.inc"
#undef SCALAR
#undef FN
#undef S_ZERO
#undef S_ACC
#undef S_SUB
#undef S_WMUL
#undef S_ADD
#undef S_MULACC

/* fp16 tables: rounding points of c10::Half's operators (Half.h:501-585) */
#define SCALAR orc_half
#define FN(n) n##_f16
#define S_ZERO ((orc_half)0.0f)
#define S_ACC(a, w, v) ((a) = f2h(h2f(a) + h2f(f2h((w) * h2f(v)))))
#define S_SUB(r, l) f2h(h2f(r) - h2f(l))
#define S_WMUL(w, v) f2h((w) * h2f(v))
#define S_ADD(a, b) f2h(h2f(a) + h2f(b))
#define S_MULACC(a, x, y) ((a) = f2h(h2f(a) + h2f(f2h(h2f(x) * h2f(y)))))
#include "grid_impl.inc"
#undef SCALAR
#undef FN
#undef S_ZERO
#undef S_ACC
#undef S_SUB
#undef S_WMUL
#undef S_ADD
#undef S_MULACC

/* fp64 tables */
#define SCALAR double
#define FN(n) n##_f64
#define S_ZERO 0.0
#define S_ACC(a, w, v) ((a) = fma((double)(w), (v), (a)))
#define S_SUB(r, l) ((r) - (l))
#define S_WMUL(w, v) ((double)(w) * (v))
#define S_ADD(a, b) ((a) + (b))
#define S_MULACC(a, x, y) ((a) = fma((x), (y), (a)))
#include "grid_impl.inc"
#undef SCALAR
#undef FN
#undef S_ZERO
#undef S_ACC
#undef S_SUB
#undef S_WMUL
#undef S_ADD
#undef S_MULACC

/* ------------------------------------------------------------------------------------------------
 *                                     fully-fused MLP (fp16)
 * ---------------------------------------------------------------------------------------------- */

/* activation ids, ffmlp.cu:22-33 */
enum { ACT_RELU = 0, ACT_EXP = 1, ACT_SINE = 2, ACT_SIGMOID = 3, ACT_SQUAREPLUS = 4, ACT_SOFTPLUS = 5, ACT_NONE = 6 };
#define K_ACT 10.0f /* utils.h: static constexpr float K_ACT = 10.0f */

/* utils.h:425-470 (forward), evaluated on the fp16-rounded pre-activation like the wmma fragment path */
static inline orc_half orc_act(int act, orc_half v)
{
    const float x = h2f(v);
    switch (act) {
        case ACT_RELU: return f2h(x * (float)(x > 0.0f));
        case ACT_EXP: return f2h(expf(x));
        case ACT_SINE: return f2h(sinf(x));
        case ACT_SIGMOID: return f2h(1.0f / (1.0f + expf(-x)));
        case ACT_SQUAREPLUS: { float t = x * K_ACT; return f2h(0.5f * (t + sqrtf(t * t + 4)) / K_ACT); }
        case ACT_SOFTPLUS: return f2h(logf(expf(x * K_ACT) + 1.0f) / K_ACT);
        default: return v;
    }
}

/* utils.h:538-588 (backward through the activation, given the stored post-activation value) */
static inline orc_half orc_act_bwd(int act, orc_half g, orc_half fwd)
{
    const float f = h2f(fwd);
    switch (act) {
        case ACT_RELU: return f2h(h2f(g) * (float)(f > 0.0f));
        case ACT_EXP: return f2h(h2f(g) * f);
        case ACT_SIGMOID: return f2h(h2f(g) * h2f(f2h(f * h2f(f2h(1.0f - f)))));
        case ACT_SQUAREPLUS: { float yv = f * K_ACT; return f2h(h2f(g) * h2f(f2h(yv * yv / (yv * yv + 1)))); }
        case ACT_SOFTPLUS: return f2h(h2f(g) * h2f(f2h(1.0f - expf(-f * K_ACT))));
        case ACT_SINE: return g; /* unsupported in the reference too (utils.h:552) */
        default: return g;
    }
}

static inline float orc_dot_h(const orc_half* a, const orc_half* w, uint32_t K, int acc_mode)
{
    if (acc_mode == 0) {
        float s = 0.0f;
        for (uint32_t k = 0; k < K; k++) s += h2f(a[k]) * h2f(w[k]);
        return s;
    }
    orc_half acc = (orc_half)0.0f; /* fp16 accumulator rounded after each 16-wide block (wmma m16n16k16, ffmlp.cu:68,101) */
    for (uint32_t k0 = 0; k0 < K; k0 += 16) {
        float s = h2f(acc);
        for (uint32_t k = k0; k < k0 + 16 && k < K; k++) s += h2f(a[k]) * h2f(w[k]);
        acc = f2h(s);
    }
    return h2f(acc);
}

/* Y = act_out(W_last . act(W_{n-1} ... act(W_0 X)))   no biases;  all matrices row-major [out,in] (ffmlp.cu:632).
 * inputs [B,in] fp16, outputs [B,out_dim] fp16 (out_dim <= 16 padded to 16 by the Python layer),
 * forward_buffer (nullable) [num_layers, B, hidden] post-activation of every hidden layer (ffmlp.cu:121-128). */
int orc_ffmlp_forward(const orc_half* inputs, const orc_half* weights, orc_half* outputs, orc_half* forward_buffer,
                      uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t num_layers,
                      int act, int out_act, int acc_mode)
{
    if (hidden != 16 && hidden != 32 && hidden != 64 && hidden != 128 && hidden != 256) return -3; /* ffmlp.cu:658 */
    #pragma omp parallel
    {
        orc_half* cur = (orc_half*)malloc(sizeof(orc_half) * 2 * 256);
        orc_half* nxt = cur + 256;
        #pragma omp for schedule(static)
        for (int64_t b = 0; b < (int64_t)B; b++) {
            const orc_half* x = inputs + (size_t)b * in_dim;
            const orc_half* W = weights;
            for (uint32_t o = 0; o < hidden; o++) nxt[o] = orc_act(act, f2h(orc_dot_h(x, W + (size_t)o * in_dim, in_dim, acc_mode)));
            W += (size_t)hidden * in_dim;
            if (forward_buffer) memcpy(forward_buffer + ((size_t)0 * B + b) * hidden, nxt, sizeof(orc_half) * hidden);
            for (uint32_t l = 1; l < num_layers; l++) {
                orc_half* t = cur; cur = nxt; nxt = t;
                for (uint32_t o = 0; o < hidden; o++) nxt[o] = orc_act(act, f2h(orc_dot_h(cur, W + (size_t)o * hidden, hidden, acc_mode)));
                W += (size_t)hidden * hidden;
                if (forward_buffer) memcpy(forward_buffer + ((size_t)l * B + b) * hidden, nxt, sizeof(orc_half) * hidden);
            }
            for (uint32_t o = 0; o < out_dim; o++)
                outputs[(size_t)b * out_dim + o] = orc_act(out_act, f2h(orc_dot_h(nxt, W + (size_t)o * hidden, hidden, acc_mode)));
        }
        free(cur < nxt ? cur : nxt);
    }
    return 0;
}

/* ffmlp_backward (ffmlp.cu:749-895) + kernel_mlp_fused_backward (:411-518), fp32 accumulation:
 *   backward_buffer[0]   = (grad . W_last) * act'(fwd[n-1])                         [B,hidden]
 *   backward_buffer[k+1] = (backward_buffer[k] . W_hidden[n-2-k]) * act'(fwd[n-2-k])
 *   grad_inputs          = backward_buffer[n-1] . W_0
 *   grad_W_last = grad^T fwd[n-1];  grad_W_hidden[j] = dpre[j+1]^T fwd[j];  grad_W_0 = dpre[0]^T X
 * where n = num_layers and dpre[j] = backward_buffer[n-1-j].  Weight gradients are accumulated in fp64 here
 * (the reference sums them in fp16 split-K GEMMs: cutlass_matmul.h:81-82) and rounded to fp16 once. */
int orc_ffmlp_backward(const orc_half* grad, const orc_half* inputs, const orc_half* weights, const orc_half* forward_buffer,
                       orc_half* backward_buffer, orc_half* grad_inputs /*nullable*/, orc_half* grad_weights,
                       uint32_t B, uint32_t in_dim, uint32_t out_dim, uint32_t hidden, uint32_t num_layers, int act)
{
    const uint32_t n = num_layers;
    const orc_half* W0 = weights;
    const orc_half* Wh = weights + (size_t)hidden * in_dim;                 /* hidden matrices 0..n-2 */
    const orc_half* Wl = Wh + (size_t)(n - 1) * hidden * hidden;            /* [out_dim, hidden] */
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)B; b++) {
        const orc_half* g = grad + (size_t)b * out_dim;
        orc_half* bb = backward_buffer + ((size_t)0 * B + b) * hidden;
        const orc_half* fw = forward_buffer + ((size_t)(n - 1) * B + b) * hidden;
        for (uint32_t i = 0; i < hidden; i++) {
            float s = 0.0f;
            for (uint32_t o = 0; o < out_dim; o++) s += h2f(g[o]) * h2f(Wl[(size_t)o * hidden + i]);
            bb[i] = orc_act_bwd(act, f2h(s), fw[i]);
        }
        for (uint32_t k = 0; k + 1 < n; k++) {
            const orc_half* Wk = Wh + (size_t)(n - 2 - k) * hidden * hidden;
            const orc_half* src = backward_buffer + ((size_t)k * B + b) * hidden;
            orc_half* dst = backward_buffer + ((size_t)(k + 1) * B + b) * hidden;
            const orc_half* fk = forward_buffer + ((size_t)(n - 2 - k) * B + b) * hidden;
            for (uint32_t i = 0; i < hidden; i++) {
                float s = 0.0f;
                for (uint32_t o = 0; o < hidden; o++) s += h2f(src[o]) * h2f(Wk[(size_t)o * hidden + i]);
                dst[i] = orc_act_bwd(act, f2h(s), fk[i]);
            }
        }
        if (grad_inputs) {
            const orc_half* src = backward_buffer + ((size_t)(n - 1) * B + b) * hidden;
            for (uint32_t i = 0; i < in_dim; i++) {
                float s = 0.0f;
                for (uint32_t o = 0; o < hidden; o++) s += h2f(src[o]) * h2f(W0[(size_t)o * in_dim + i]);
                grad_inputs[(size_t)b * in_dim + i] = f2h(s);
            }
        }
    }
    /* weight gradients */
    orc_half* gW0 = grad_weights;
    orc_half* gWh = grad_weights + (size_t)hidden * in_dim;
    orc_half* gWl = gWh + (size_t)(n - 1) * hidden * hidden;
    #pragma omp parallel for schedule(static)
    for (int64_t oi = 0; oi < (int64_t)out_dim * hidden; oi++) {
        const uint32_t o = (uint32_t)(oi / hidden), i = (uint32_t)(oi % hidden);
        double s = 0;
        for (uint32_t b = 0; b < B; b++)
            s += (double)h2f(grad[(size_t)b * out_dim + o]) * (double)h2f(forward_buffer[((size_t)(n - 1) * B + b) * hidden + i]);
        gWl[oi] = f2h((float)s);
    }
    for (uint32_t j = 0; j + 1 < n; j++) {
        const orc_half* dpre = backward_buffer + (size_t)(n - 2 - j) * B * hidden; /* dpre[j+1] */
        const orc_half* fj = forward_buffer + (size_t)j * B * hidden;
        #pragma omp parallel for schedule(static)
        for (int64_t oi = 0; oi < (int64_t)hidden * hidden; oi++) {
            const uint32_t o = (uint32_t)(oi / hidden), i = (uint32_t)(oi % hidden);
            double s = 0;
            for (uint32_t b = 0; b < B; b++) s += (double)h2f(dpre[(size_t)b * hidden + o]) * (double)h2f(fj[(size_t)b * hidden + i]);
            gWh[(size_t)j * hidden * hidden + oi] = f2h((float)s);
        }
    }
    {
        const orc_half* dpre0 = backward_buffer + (size_t)(n - 1) * B * hidden;
        #pragma omp parallel for schedule(static)
        for (int64_t oi = 0; oi < (int64_t)hidden * in_dim; oi++) {
            const uint32_t o = (uint32_t)(oi / in_dim), i = (uint32_t)(oi % in_dim);
            double s = 0;
            for (uint32_t b = 0; b < B; b++) s += (double)h2f(dpre0[(size_t)b * hidden + o]) * (double)h2f(inputs[(size_t)b * in_dim + i]);
            gW0[oi] = f2h((float)s);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 *                                   spherical-harmonics encoder
 * ---------------------------------------------------------------------------------------------- */
#include "sh_table.inc"

/* kernel_sh (shencoder.cu:28-356): outputs [B, C*C]; dy_dx [B, 3, C*C] (dx block, dy block, dz block).
 * Evaluated from the monomial table in fp64 and rounded once to fp32 (the GPU evaluates in fp32). */
int orc_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t C, int calc_grad_inputs, float* dy_dx)
{
    if (D != 3 || C < 1 || C > 8) return -4;
    const uint32_t C2 = C * C;
    #pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)B; b++) {
        double px[8], py[8], pz[8], acc[4][64];
        px[0] = py[0] = pz[0] = 1.0;
        for (int i = 1; i < 8; i++) {
            px[i] = px[i - 1] * (double)inputs[b * 3 + 0];
            py[i] = py[i - 1] * (double)inputs[b * 3 + 1];
            pz[i] = pz[i - 1] * (double)inputs[b * 3 + 2];
        }
        memset(acc, 0, sizeof(acc));
        for (int t = 0; t < ORC_SH_NTERMS; t++) {
            const orc_sh_term_t* e = &ORC_SH_TERMS[t];
            if (e->out >= C2) continue;
            acc[e->kind][e->out] += e->coeff * px[e->a] * py[e->b] * pz[e->c];
        }
        for (uint32_t k = 0; k < C2; k++) outputs[(size_t)b * C2 + k] = (float)acc[0][k];
        if (calc_grad_inputs)
            for (uint32_t d = 0; d < 3; d++)
                for (uint32_t k = 0; k < C2; k++) dy_dx[((size_t)b * 3 + d) * C2 + k] = (float)acc[1 + d][k];
    }
    return 0;
}

/* kernel_sh_backward (shencoder.cu:360-381): grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch] */
int orc_sh_encode_backward(const float* grad, uint32_t B, uint32_t D, uint32_t C, const float* dy_dx, float* grad_inputs)
{
    const uint32_t C2 = C * C;
    #pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)B * D; t++) {
        const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (int64_t)b * D);
        float r = grad_inputs[t];
        for (uint32_t ch = 0; ch < C2; ch++) r = fmaf(grad[(size_t)b * C2 + ch], dy_dx[((size_t)b * D + d) * C2 + ch], r);
        grad_inputs[t] = r;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 *                                          ray marching
 * ---------------------------------------------------------------------------------------------- */

#define ORC_SQRT3 1.7320508075688772f
#define ORC_RPI 0.3183098861837907f

static inline float orc_clamp(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }
static inline float orc_sign(float x) { return copysignf(1.0f, x); }

/* raymarching.cu:44-56 */
static inline int orc_mip_from_pos(float x, float y, float z, float max_cascade)
{
    const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
    int e; frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}
static inline int orc_mip_from_dt(float dt, float H, float max_cascade)
{
    const float mx = (float)((double)(dt * H) * 0.5);
    int e; frexpf(mx, &e);
    return (int)fminf(max_cascade - 1, fmaxf(0, (float)e));
}

/* raymarching.cu:58-83 */
static inline uint32_t orc_expand_bits(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
static inline uint32_t orc_morton3D_1(uint32_t x, uint32_t y, uint32_t z)
{
    return orc_expand_bits(x) | (orc_expand_bits(y) << 1) | (orc_expand_bits(z) << 2);
}
static inline uint32_t orc_morton3D_invert_1(uint32_t x)
{
    x = x & 0x49249249;
    x = (x | (x >> 2)) & 0xc30c30c3;
    x = (x | (x >> 4)) & 0x0f00f00f;
    x = (x | (x >> 8)) & 0xff0000ff;
    x = (x | (x >> 16)) & 0x0000ffff;
    return x;
}

void orc_morton3D(const int* coords, uint32_t N, int* indices)
{
    for (uint32_t n = 0; n < N; n++) indices[n] = (int)orc_morton3D_1((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}
void orc_morton3D_invert(const int* indices, uint32_t N, int* coords)
{
    for (uint32_t n = 0; n < N; n++) {
        const int ind = indices[n];
        coords[n * 3 + 0] = (int)orc_morton3D_invert_1((uint32_t)(ind >> 0));
        coords[n * 3 + 1] = (int)orc_morton3D_invert_1((uint32_t)(ind >> 1));
        coords[n * 3 + 2] = (int)orc_morton3D_invert_1((uint32_t)(ind >> 2));
    }
}

/* raymarching.cu:270-291 */
void orc_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield)
{
    #pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        uint8_t bits = 0;
        for (int i = 0; i < 8; i++) bits |= (grid[n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
        bitfield[n] = bits;
    }
}

/* raymarching.cu:94-147 */
void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near, float* nears, float* fars)
{
    #pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float rdx = 1 / rays_d[n * 3], rdy = 1 / rays_d[n * 3 + 1], rdz = 1 / rays_d[n * 3 + 2];
        float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx, t;
        if (near > far) { t = near; near = far; far = t; }
        float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
        if (near_y > far_y) { t = near_y; near_y = far_y; far_y = t; }
        if (near > far_y || near_y > far) { nears[n] = fars[n] = 3.402823466e+38f; continue; }
        if (near_y > near) near = near_y;
        if (far_y < far) far = far_y;
        float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
        if (near_z > far_z) { t = near_z; near_z = far_z; far_z = t; }
        if (near > far_z || near_z > far) { nears[n] = fars[n] = 3.402823466e+38f; continue; }
        if (near_z > near) near = near_z;
        if (far_z < far) far = far_z;
        if (near < min_near) near = min_near;
        nears[n] = near; fars[n] = far;
    }
}

/* raymarching.cu:165-200 */
void orc_polar_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords)
{
    #pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float A = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float Bq = fmaf(oz, dz, fmaf(oy, dy, ox * dx));
        const float Cq = fmaf(oz, oz, fmaf(oy, oy, ox * ox)) - radius * radius;
        const float t = (-Bq + sqrtf(Bq * Bq - A * Cq)) / A;
        const float x = fmaf(t, dx, ox), y = fmaf(t, dy, oy), z = fmaf(t, dz, oz);
        const float theta = atan2f(sqrtf(fmaf(z, z, x * x)), y);
        const float phi = atan2f(z, x);
        coords[n * 2 + 0] = 2 * theta * ORC_RPI - 1;
        coords[n * 2 + 1] = phi * ORC_RPI;
    }
}

/* pcg32.h:44-170 */
typedef struct { uint64_t state, inc; } orc_pcg32;
#define ORC_PCG32_MULT 0x5851f42d4c957f2dULL
static inline uint32_t orc_pcg_next_uint(orc_pcg32* r)
{
    uint64_t old = r->state;
    r->state = old * ORC_PCG32_MULT + r->inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
static inline void orc_pcg_seed(orc_pcg32* r, uint64_t initstate, uint64_t initseq)
{
    r->state = 0U; r->inc = (initseq << 1u) | 1u;
    orc_pcg_next_uint(r); r->state += initstate; orc_pcg_next_uint(r);
}
static inline float orc_pcg_next_float(orc_pcg32* r)
{
    union { uint32_t u; float f; } x;
    x.u = (orc_pcg_next_uint(r) >> 9) | 0x3f800000u;
    return x.f - 1.0f;
}
static inline void orc_pcg_advance(orc_pcg32* r, int64_t delta_)
{
    uint64_t cur_mult = ORC_PCG32_MULT, cur_plus = r->inc, acc_mult = 1u, acc_plus = 0u, delta = (uint64_t)delta_;
    while (delta > 0) {
        if (delta & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta /= 2;
    }
    r->state = acc_mult * r->state + acc_plus;
}
/* known-answer helper: stream of `count` next_uint() after seeding with (seed, 1) and advance(adv) */
void orc_pcg32_stream(uint64_t seed, int64_t adv, uint32_t count, uint32_t* out_u, float* out_f)
{
    orc_pcg32 r; orc_pcg_seed(&r, seed, 1);
    if (adv) orc_pcg_advance(&r, adv);
    orc_pcg32 r2 = r;
    for (uint32_t i = 0; i < count; i++) { out_u[i] = orc_pcg_next_uint(&r); if (out_f) out_f[i] = orc_pcg_next_float(&r2); }
}

/* One DDA probe, shared by the three marchers (raymarching.cu:362-402 / 430-481 / 951-1004).
 * Returns 1 if the voxel at parameter t is occupied (and dt_out = step), else advances *t past the voxel. */
typedef struct {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, rH, bound, dt_gamma, dt_min, dt_max;
    uint32_t C, H; const uint8_t* grid;
} orc_ray_ctx;

static inline int orc_probe(const orc_ray_ctx* c, float* t, float* x, float* y, float* z, float* dt_out)
{
    *x = orc_clamp(fmaf(*t, c->dx, c->ox), -c->bound, c->bound);
    *y = orc_clamp(fmaf(*t, c->dy, c->oy), -c->bound, c->bound);
    *z = orc_clamp(fmaf(*t, c->dz, c->oz), -c->bound, c->bound);
    const float dt = orc_clamp(*t * c->dt_gamma, c->dt_min, c->dt_max);
    const int m1 = orc_mip_from_pos(*x, *y, *z, (float)c->C), m2 = orc_mip_from_dt(dt, (float)c->H, (float)c->C);
    const int level = m1 > m2 ? m1 : m2;
    const float mip_bound = fminf((float)(1 << level), c->bound);
    const float mip_rbound = 1 / mip_bound;
    const float Hm1 = (float)(c->H - 1);
    const int nx = (int)orc_clamp((float)(0.5 * (double)fmaf(*x, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const int ny = (int)orc_clamp((float)(0.5 * (double)fmaf(*y, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const int nz = (int)orc_clamp((float)(0.5 * (double)fmaf(*z, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
    const uint32_t index = (uint32_t)level * c->H * c->H * c->H + orc_morton3D_1((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
    const int occ = c->grid[index / 8] & (1 << (index % 8));
    *dt_out = dt;
    if (occ) return 1;
    const float tx = (fmaf(((float)nx + 0.5f + 0.5f * orc_sign(c->dx)) * c->rH * 2 - 1, mip_bound, -*x)) * c->rdx;
    const float ty = (fmaf(((float)ny + 0.5f + 0.5f * orc_sign(c->dy)) * c->rH * 2 - 1, mip_bound, -*y)) * c->rdy;
    const float tz = (fmaf(((float)nz + 0.5f + 0.5f * orc_sign(c->dz)) * c->rH * 2 - 1, mip_bound, -*z)) * c->rdz;
    const float tt = *t + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
    do { *t += orc_clamp(*t * c->dt_gamma, c->dt_min, c->dt_max); } while (*t < tt);
    return 0;
}

static inline void orc_ray_ctx_init(orc_ray_ctx* c, const float* o, const float* d, float bound, float dt_gamma,
                                    uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid)
{
    c->ox = o[0]; c->oy = o[1]; c->oz = o[2]; c->dx = d[0]; c->dy = d[1]; c->dz = d[2];
    c->rdx = 1 / c->dx; c->rdy = 1 / c->dy; c->rdz = 1 / c->dz; c->rH = 1 / (float)H;
    c->bound = bound; c->dt_gamma = dt_gamma;
    c->dt_min = 2 * ORC_SQRT3 / (float)max_steps;
    c->dt_max = 2 * ORC_SQRT3 * (float)(1 << (C - 1)) / (float)H;
    c->C = C; c->H = H; c->grid = grid;
}

/* kernel_march_rays_train (raymarching.cu:314-483), rays visited in ascending n (a deterministic member of the
 * reference's atomicAdd outcome set).  rays_ts (nullable) is the _differentiable variant's extra output (:658). */
void orc_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                          uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                          const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas, float* rays_ts,
                          int* rays, int* counter, uint32_t perturb)
{
    for (uint32_t n = 0; n < N; n++) {
        orc_ray_ctx c; orc_ray_ctx_init(&c, rays_o + n * 3, rays_d + n * 3, bound, dt_gamma, max_steps, C, H, grid);
        const float far = fars[n];
        float t0 = nears[n];
        if (perturb) { orc_pcg32 r; orc_pcg_seed(&r, 42, 1); orc_pcg_advance(&r, (int64_t)n); t0 = fmaf(c.dt_min, orc_pcg_next_float(&r), t0); /* contracted on the GPU */ }
        float t = t0, x, y, z, dt; uint32_t num_steps = 0;
        while (t < far && num_steps < max_steps) { if (orc_probe(&c, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; } }
        const uint32_t point_index = (uint32_t)counter[0]; counter[0] += (int)num_steps;
        const uint32_t ray_index = (uint32_t)counter[1]; counter[1] += 1;
        rays[ray_index * 3] = (int)n; rays[ray_index * 3 + 1] = (int)point_index; rays[ray_index * 3 + 2] = (int)num_steps;
        if (num_steps == 0) continue;
        if (point_index + num_steps >= M) continue;
        float* px = xyzs + (size_t)point_index * 3; float* pd = dirs + (size_t)point_index * 3; float* pl = deltas + (size_t)point_index * 2;
        float* pt = rays_ts ? rays_ts + point_index : NULL;
        t = t0; uint32_t step = 0; float last_t = t;
        while (t < far && step < num_steps) {
            if (orc_probe(&c, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z; pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt; pl[0] = dt; pl[1] = t - last_t; if (pt) { pt[0] = t; pt++; } last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* kernel_composite_rays_train_forward (raymarching.cu:700-777) */
void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int* rays,
                                      uint32_t M, uint32_t N, float* weights_sum, float* depth, float* image)
{
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps >= M) {
            weights_sum[index] = 0; depth[index] = 0; image[index * 3] = image[index * 3 + 1] = image[index * 3 + 2] = 0;
            continue;
        }
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, t = 0, d = 0;
        for (uint32_t s = 0; s < num_steps; s++) {
            const size_t i = offset + s;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[i * 2]);
            const float weight = alpha * T;
            r = fmaf(weight, rgbs[i * 3], r); g = fmaf(weight, rgbs[i * 3 + 1], g); b = fmaf(weight, rgbs[i * 3 + 2], b);
            t += deltas[i * 2 + 1];
            d = fmaf(weight, t, d);
            ws += weight;
            T *= 1.0f - alpha;
        }
        weights_sum[index] = ws; depth[index] = d; image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* kernel_composite_rays_train_backward (raymarching.cu:802-881); grads pre-zeroed by the caller (raymarching.py:334-335) */
void orc_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas, const float* rgbs,
                                       const float* deltas, const int* rays, const float* weights_sum, const float* image,
                                       uint32_t M, uint32_t N, float* grad_sigmas, float* grad_rgbs)
{
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)N; n++) {
        const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
        if (num_steps == 0 || offset + num_steps >= M) continue;
        const float gws = grad_weights_sum[index];
        const float* gi = grad_image + (size_t)index * 3;
        const float r_final = image[index * 3], g_final = image[index * 3 + 1], b_final = image[index * 3 + 2], ws_final = weights_sum[index];
        float T = 1.0f, r = 0, g = 0, b = 0, ws = 0;
        for (uint32_t s = 0; s < num_steps; s++) {
            const size_t i = offset + s;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[i * 2]);
            const float weight = alpha * T;
            r = fmaf(weight, rgbs[i * 3], r); g = fmaf(weight, rgbs[i * 3 + 1], g); b = fmaf(weight, rgbs[i * 3 + 2], b);
            ws += weight;
            T *= 1.0f - alpha;
            grad_rgbs[i * 3] = gi[0] * weight; grad_rgbs[i * 3 + 1] = gi[1] * weight; grad_rgbs[i * 3 + 2] = gi[2] * weight;
            grad_sigmas[i] = deltas[i * 2] * (
                gi[0] * (T * rgbs[i * 3] - (r_final - r)) +
                gi[1] * (T * rgbs[i * 3 + 1] - (g_final - g)) +
                gi[2] * (T * rgbs[i * 3 + 2] - (b_final - b)) +
                gws * (T - (ws_final - ws)));
        }
    }
}

/* kernel_march_rays (raymarching.cu:900-1006); outputs pre-zeroed by the caller (raymarching.py:389-391) */
void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, const float* rays_t, const float* rays_o, const float* rays_d,
                    float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                    const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas, uint32_t perturb)
{
    #pragma omp parallel for schedule(dynamic, 64)
    for (int64_t n = 0; n < (int64_t)n_alive; n++) {
        const int index = rays_alive[n];
        float t = rays_t[n];
        orc_ray_ctx c; orc_ray_ctx_init(&c, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, bound, dt_gamma, max_steps, C, H, grid);
        const float far = fars[index];
        float* px = xyzs + (size_t)n * n_step * 3; float* pd = dirs + (size_t)n * n_step * 3; float* pl = deltas + (size_t)n * n_step * 2;
        if (perturb) { orc_pcg32 r; orc_pcg_seed(&r, (uint64_t)perturb, 1); orc_pcg_advance(&r, (int64_t)n); t = fmaf(c.dt_min, orc_pcg_next_float(&r), t); /* contracted on the GPU */ }
        float last_t = t, x, y, z, dt; uint32_t step = 0;
        while (t < far && step < n_step) {
            if (orc_probe(&c, &t, &x, &y, &z, &dt)) {
                px[0] = x; px[1] = y; px[2] = z; pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
                t += dt; pl[0] = dt; pl[1] = t - last_t; last_t = t;
                px += 3; pd += 3; pl += 2; step++;
            }
        }
    }
}

/* kernel_composite_rays (raymarching.cu:1021-1104): in-place continuation of the accumulation */
void orc_composite_rays(uint32_t n_alive, uint32_t n_step, const int* rays_alive, float* rays_t, const float* sigmas, const float* rgbs,
                        const float* deltas, float* weights_sum, float* depth, float* image)
{
    #pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < (int64_t)n_alive; n++) {
        const int index = rays_alive[n];
        float t = rays_t[n];
        float weight_sum = weights_sum[index], d = depth[index], r = image[index * 3], g = image[index * 3 + 1], b = image[index * 3 + 2];
        uint32_t step = 0;
        while (step < n_step) {
            const size_t i = (size_t)n * n_step + step;
            if (deltas[i * 2] == 0) break;
            const float alpha = 1.0f - expf(-sigmas[i] * deltas[i * 2]);
            const float T = 1 - weight_sum;
            const float weight = alpha * T;
            weight_sum += weight;
            t += deltas[i * 2 + 1];
            d = fmaf(weight, t, d);
            r = fmaf(weight, rgbs[i * 3], r); g = fmaf(weight, rgbs[i * 3 + 1], g); b = fmaf(weight, rgbs[i * 3 + 2], b);
            if ((double)T < 1e-4) break;   /* :1081 compares against a double literal */
            step++;
        }
        rays_t[n] = (step < n_step) ? -1.0f : t;
        weights_sum[index] = weight_sum; depth[index] = d; image[index * 3] = r; image[index * 3 + 1] = g; image[index * 3 + 2] = b;
    }
}

/* kernel_compact_rays (raymarching.cu:1117-1134) — survivors kept in ascending slot order */
void orc_compact_rays(uint32_t n_alive, int* rays_alive, const int* rays_alive_old, float* rays_t, const float* rays_t_old, int* alive_counter)
{
    for (uint32_t n = 0; n < n_alive; n++) {
        if (rays_t_old[n] >= 0) {
            const int index = alive_counter[0]++;
            rays_alive[index] = rays_alive_old[n];
            rays_t[index] = rays_t_old[n];
        }
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
