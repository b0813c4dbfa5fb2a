/*
 * TEST INFRASTRUCTURE ONLY -- see oracle_common.h.
 *
 * CPU restatement of the grid regularisers of /root/reference/svox2/csrc/loss_kernel.cu (cited per function) and
 * add_surface_normal_grad (include/render_util.cuh:1824-2133).  One loop iteration == one CUDA thread of the reference.
 * Pinning: the reference ships no vectors and no CPU implementation of these; tests/test_loss_gpu.py checks this file
 * against the UNMODIFIED reference kernels (oracle/_ref) on the GPU box, and the CUDA port against both.
 */
#include "oracle_common.h"

#define OMP_MIN_CELLS 200000
#define ATOMIC_ADD(dst, val) do { const float _v = (val); _Pragma("omp atomic") (dst) += _v; } while (0)

static void ray_scale(const int32_t *size, float *s) { /* :22-62 */
    s[0] = size[0] * (1.f / 256.f);
    s[1] = size[1] * (1.f / 256.f);
    s[2] = size[2] * (1.f / 256.f);
}
#define LNK(x, y, z) links[((int64_t)(x) * size[1] + (y)) * size[2] + (z)]

/* tv_kernel :72-117 + host :1214-1247 ; accumulates in double (the reference's block/atomic order is unspecified) */
float oracle_tv(const int32_t *links, const int32_t *size, const float *data, int n_cols, int start_dim, int end_dim,
                int ignore_edge) {
    float sc[3];
    ray_scale(size, sc);
    const int64_t nl = (int64_t)(size[0] - 1) * (size[1] - 1) * (size[2] - 1);
    double acc = 0;
    for (int x = 0; x < size[0] - 1; ++x)
        for (int y = 0; y < size[1] - 1; ++y)
            for (int z = 0; z < size[2] - 1; ++z)
                for (int idx = start_dim; idx < end_dim; ++idx) {
                    if (ignore_edge && LNK(x, y, z) == 0) continue;
                    const float v000 = LNK(x, y, z) >= 0 ? data[(int64_t)LNK(x, y, z) * n_cols + idx] : 0.f;
                    const float nullv = ignore_edge ? v000 : 0.f;
                    const float v100 = LNK(x + 1, y, z) >= 0 ? data[(int64_t)LNK(x + 1, y, z) * n_cols + idx] : nullv;
                    const float v010 = LNK(x, y + 1, z) >= 0 ? data[(int64_t)LNK(x, y + 1, z) * n_cols + idx] : nullv;
                    const float v001 = LNK(x, y, z + 1) >= 0 ? data[(int64_t)LNK(x, y, z + 1) * n_cols + idx] : nullv;
                    const float dx = (v100 - v000) * sc[0], dy = (v010 - v000) * sc[1], dz = (v001 - v000) * sc[2];
                    acc += sqrtf(1e-5f + dx * dx + dy * dy + dz * dz);
                }
    return (float)(acc * (double)(1.f / (float)nl));
}

/* tv_grad_kernel :119-184 + host :1249-1287 */
void oracle_tv_grad(const int32_t *links, const int32_t *size, const float *data, int n_cols, int start_dim, int end_dim,
                    float scale, int ignore_edge, float *grad) {
    float sc[3];
    ray_scale(size, sc);
    const int64_t nl = (int64_t)(size[0] - 1) * (size[1] - 1) * (size[2] - 1);
    scale = scale / (float)nl;
    for (int x = 0; x < size[0] - 1; ++x)
        for (int y = 0; y < size[1] - 1; ++y)
            for (int z = 0; z < size[2] - 1; ++z)
                for (int idx = start_dim; idx < end_dim; ++idx) {
                    const int32_t l000 = LNK(x, y, z), l100 = LNK(x + 1, y, z), l010 = LNK(x, y + 1, z), l001 = LNK(x, y, z + 1);
                    if (ignore_edge && l000 == 0) continue;
                    float v000 = 0.f, v100 = 0.f, v010 = 0.f, v001 = 0.f;
                    if (l000 >= 0) v000 = data[(int64_t)l000 * n_cols + idx];
                    if (l100 >= 0) v100 = data[(int64_t)l100 * n_cols + idx]; else if (ignore_edge) v100 = v000;
                    if (l010 >= 0) v010 = data[(int64_t)l010 * n_cols + idx]; else if (ignore_edge) v010 = v000;
                    if (l001 >= 0) v001 = data[(int64_t)l001 * n_cols + idx]; else if (ignore_edge) v001 = v000;
                    float dx = v100 - v000, dy = v010 - v000, dz = v001 - v000;
                    const float idelta = scale * (1.f / sqrtf(1e-9f + dx * dx + dy * dy + dz * dz));
                    dx *= sc[0]; dy *= sc[1]; dz *= sc[2];
                    if (dx != 0.f && l100 >= 0) grad[(int64_t)l100 * n_cols + idx] += dx * idelta;
                    if (dy != 0.f && l010 >= 0) grad[(int64_t)l010 * n_cols + idx] += dy * idelta;
                    if (dz != 0.f && l001 >= 0) grad[(int64_t)l001 * n_cols + idx] += dz * idelta;
                    if (l000 >= 0) grad[(int64_t)l000 * n_cols + idx] += -(dx + dy + dz) * idelta;
                }
}

/* tv_grad_sparse_kernel :738-807 (surf == 0) and surf_tv_grad_sparse_kernel :809-893 (surf == 1) + hosts :1327-1427 */
void oracle_tv_grad_sparse(const int32_t *links, const int32_t *size, const float *data, int n_cols, const float *density,
                           int density_cols, const int32_t *cells, int64_t n_cells, uint8_t *mask, int start_dim,
                           int end_dim, float scale, int ignore_edge, float edge_value, int ignore_last_z,
                           int alpha_dependency, int surf, float *grad) {
    float sc[3];
    ray_scale(size, sc);
    scale = scale / (float)(int)n_cells;
    const int64_t offx = (int64_t)size[1] * size[2];
    const int offy = size[2];
    /* lists above OMP_MIN_CELLS entries (bench.py's CPU arm) run on all host threads, like the CUDA grid; the accumulations
     * are atomic then and their order is unspecified, as on the GPU.  The parity tests stay below it: serial, one order. */
#pragma omp parallel for schedule(static) if (n_cells > OMP_MIN_CELLS)
    for (int64_t i = 0; i < n_cells; ++i)
        for (int idx = start_dim; idx < end_dim; ++idx) {
            const int64_t xyz = cells[i];
            const int z = (int)(xyz % size[2]);
            const int64_t xy = xyz / size[2];
            const int y = (int)(xy % size[1]);
            const int x = (int)(xy / size[1]);
            const int32_t *lp = links + xyz;
            if (ignore_edge && *lp == 0) continue;
            const int32_t l000 = lp[0];
            const int32_t l001 = ((z + 1 < size[2]) && (!ignore_last_z || z != size[2] - 2)) ? lp[1] : 0;
            const int32_t l010 = (y + 1 < size[1]) ? lp[offy] : 0;
            const int32_t l100 = (x + 1 < size[0]) ? lp[offx] : 0;
            if (ignore_last_z && z == size[2] - 2) continue;
            const float missing = surf ? edge_value : 0.f;
            const float v000 = l000 >= 0 ? data[(int64_t)l000 * n_cols + idx] : missing;
            const float nullv = ignore_edge ? v000 : missing;
            const float v001 = l001 >= 0 ? data[(int64_t)l001 * n_cols + idx] : nullv;
            const float v010 = l010 >= 0 ? data[(int64_t)l010 * n_cols + idx] : nullv;
            const float v100 = l100 >= 0 ? data[(int64_t)l100 * n_cols + idx] : nullv;
            float dx = v100 - v000, dy = v010 - v000, dz = v001 - v000;
            float idelta = scale * (1.f / sqrtf(1e-9f + dx * dx + dy * dy + dz * dz));
            if (surf && alpha_dependency) {
                const float a000 = l000 >= 0 ? density[(int64_t)l000 * density_cols + idx] : 0.f;
                const float a001 = l001 >= 0 ? density[(int64_t)l001 * density_cols + idx] : 0.f;
                const float a010 = l010 >= 0 ? density[(int64_t)l010 * density_cols + idx] : 0.f;
                const float a100 = l100 >= 0 ? density[(int64_t)l100 * density_cols + idx] : 0.f;
                const float max_alpha = o_maxf(a000, o_maxf(a001, o_maxf(a010, a100)));
                if ((double)max_alpha < 0.1) {
                    const double den = (double)(max_alpha * 10) > 1e-1 ? (double)(max_alpha * 10) : 1e-1;
                    idelta = (float)((double)idelta / den);
                }
            }
            dx *= sc[0]; dy *= sc[1]; dz *= sc[2];
            const float sm = -(dx + dy + dz);
#define MAYBE(l, v) if ((l) >= 0 && (v) != 0.f) { ATOMIC_ADD(grad[(int64_t)(l) * n_cols + idx], (v) * idelta); if (mask) mask[l] = 1; }
            MAYBE(l000, sm);
            MAYBE(l001, dz);
            MAYBE(l010, dy);
            MAYBE(l100, dx);
#undef MAYBE
        }
}

/* alpha_surf_sparsify_grad_sparse_kernel :664-734 */
void oracle_alpha_surf_sparsify(const int32_t *links, const float *alpha, int alpha_cols, const float *surf,
                                int surf_cols, const int32_t *cells, int64_t n_cells, uint8_t *mask, float scale_alpha,
                                float scale_surf, int surf_decrease, float surf_thresh, float alpha_bound,
                                float surf_bound, float *grad_alpha, float *grad_surf) {
#pragma omp parallel for schedule(static) if (n_cells > OMP_MIN_CELLS)
    for (int64_t i = 0; i < n_cells; ++i) {
        const int32_t l = links[cells[i]];
        if (l < 0) continue;
        if (mask) mask[l] = 1;
        const float a = alpha[(int64_t)l * alpha_cols];
        const float safe_grad = 1.f / o_maxf(a, 1e-8f);
        if (a > alpha_bound) ATOMIC_ADD(grad_alpha[(int64_t)l * alpha_cols], scale_alpha * safe_grad);
        const float s = surf[(int64_t)l * surf_cols];
        const int reg = surf_decrease ? (s > surf_bound) : (s < surf_bound);
        if (reg && (a < surf_thresh)) ATOMIC_ADD(grad_surf[(int64_t)l * alpha_cols], surf_decrease ? (scale_surf * safe_grad) : (-scale_surf * safe_grad));
    }
}

/* ---- add_surface_normal_grad, render_util.cuh:1870-2133 ---- */
typedef struct { int32_t l[8]; float s[8]; } Cell8;
static int load_cell(const int32_t *links, const float *surf, const int32_t *size, int x, int y, int z, Cell8 *c) {
    if (!((x < size[0] - 1) && (y < size[1] - 1) && (z < size[2] - 1))) return 0;
    for (int k = 0; k < 8; ++k) {
        c->l[k] = LNK(x + (k >> 2), y + ((k >> 1) & 1), z + (k & 1));
        if (c->l[k] < 0) return 0;
    }
    for (int k = 0; k < 8; ++k) c->s[k] = surf[c->l[k]];
    return 1;
}
static int cell_empty(const Cell8 *c, float lv) {
    int le = 1, ge = 1;
    for (int k = 0; k < 8; ++k) { le &= (c->s[k] <= lv); ge &= (c->s[k] >= lv); }
    return le || ge;
}
static void cell_normal(const Cell8 *c, float *n) {
    const float *s = c->s;
    n[0] = ((s[4] + s[5] + s[6] + s[7]) - (s[0] + s[1] + s[2] + s[3])) / 4;
    n[1] = ((s[2] + s[3] + s[6] + s[7]) - (s[0] + s[1] + s[4] + s[5])) / 4;
    n[2] = ((s[1] + s[3] + s[5] + s[7]) - (s[0] + s[2] + s[4] + s[6])) / 4;
}
static int face_connected(float s0, float s1, float s2, float s3, float lv) {
    return !(((s0 <= lv) && (s1 <= lv) && (s2 <= lv) && (s3 <= lv)) || ((s0 >= lv) && (s1 >= lv) && (s2 >= lv) && (s3 >= lv)));
}
static void scatter_normal(const Cell8 *c, const float *g, float scale, uint8_t *mask, float *grad) { /* :1824-1868 */
    for (int k = 0; k < 8; ++k) {
        const float sx = (k & 4) ? 0.25f : -0.25f, sy = (k & 2) ? 0.25f : -0.25f, sz = (k & 1) ? 0.25f : -0.25f;
        const float val = scale * (sx * g[0] + sy * g[1] + sz * g[2]);
        if (val != 0.f) { ATOMIC_ADD(grad[c->l[k]], val); if (mask) mask[c->l[k]] = 1; }
    }
}
#define NORM3(v) sqrtf(1e-9f + (v)[0] * (v)[0] + (v)[1] * (v)[1] + (v)[2] * (v)[2])
#define CUB(x) ((x) * (x) * (x))
#define SQR(x) ((x) * (x))

/* squared-difference form of the pair gradient (render_util.cuh:2042-2080; the same expressions in loss_kernel.cu:343-380) */
static void pair_l2(const float *n0, float N0, const float *n1, float N1, float *d0, float *d1) {
    const float e0 = n0[0] / N0 - n1[0] / N1, e1 = n0[1] / N0 - n1[1] / N1, e2 = n0[2] / N0 - n1[2] / N1;
    d0[0] = e0 * (-2.f * SQR(n0[0]) / CUB(N0) + 2.f / N0) + -2.f * n0[0] * n0[1] * e1 / CUB(N0) + -2.f * n0[0] * n0[2] * e2 / CUB(N0);
    d0[1] = e1 * (-2.f * SQR(n0[1]) / CUB(N0) + 2.f / N0) + -2.f * n0[0] * n0[1] * e0 / CUB(N0) + -2.f * n0[1] * n0[2] * e2 / CUB(N0);
    d0[2] = e2 * (-2.f * SQR(n0[2]) / CUB(N0) + 2.f / N0) + -2.f * n0[0] * n0[2] * e0 / CUB(N0) + -2.f * n0[1] * n0[2] * e1 / CUB(N0);
    d1[0] = e0 * (2.f * SQR(n1[0]) / CUB(N1) - 2.f / N1) + 2.f * n1[0] * n1[1] * e1 / CUB(N1) + 2.f * n1[0] * n1[2] * e2 / CUB(N1);
    d1[1] = e1 * (2.f * SQR(n1[1]) / CUB(N1) - 2.f / N1) + 2.f * n1[0] * n1[1] * e0 / CUB(N1) + 2.f * n1[1] * n1[2] * e2 / CUB(N1);
    d1[2] = e2 * (2.f * SQR(n1[2]) / CUB(N1) - 2.f / N1) + 2.f * n1[0] * n1[2] * e0 / CUB(N1) + 2.f * n1[1] * n1[2] * e1 / CUB(N1);
}

/* surface_normal_grad_sparse_kernel :397-441 + host :1572-1622 */
void oracle_surface_normal_grad_sparse(const int32_t *links, const int32_t *size, const float *surf, const int32_t *cells,
                                       int64_t n_cells, uint8_t *mask, float lv, int start_dim, int end_dim, float scale,
                                       int con_check, int ignore_empty, int use_l1, float *grad) {
    scale = scale / (float)(int)n_cells;
#pragma omp parallel for schedule(static) if (n_cells > OMP_MIN_CELLS)
    for (int64_t i = 0; i < n_cells; ++i)
        for (int rep = start_dim; rep < end_dim; ++rep) {
            const int64_t xyz = cells[i];
            const int z = (int)(xyz % size[2]);
            const int64_t xy = xyz / size[2];
            const int y = (int)(xy % size[1]);
            const int x = (int)(xy / size[1]);
            Cell8 c0, cn[3];
            if (!load_cell(links, surf, size, x, y, z, &c0)) continue;
            const int empty000 = ignore_empty ? cell_empty(&c0, lv) : 0;
            float n0[3];
            cell_normal(&c0, n0);
            int use[3];
            use[2] = load_cell(links, surf, size, x, y, z + 1, &cn[2]) &&
                     (!con_check || face_connected(c0.s[1], c0.s[3], c0.s[5], c0.s[7], lv)) &&
                     (!ignore_empty || (!empty000 || !cell_empty(&cn[2], lv)));
            use[1] = load_cell(links, surf, size, x, y + 1, z, &cn[1]) &&
                     (!con_check || face_connected(c0.s[2], c0.s[3], c0.s[6], c0.s[7], lv)) &&
                     (!ignore_empty || (!empty000 || !cell_empty(&cn[1], lv)));
            use[0] = load_cell(links, surf, size, x + 1, y, z, &cn[0]) &&
                     (!con_check || face_connected(c0.s[4], c0.s[5], c0.s[6], c0.s[7], lv)) &&
                     (!ignore_empty || (!empty000 || !cell_empty(&cn[0], lv)));
            const int norm_count = use[0] + use[1] + use[2];
            const float N0 = NORM3(n0);
            for (int a = 0; a < 3; ++a) {
                if (!use[a]) continue;
                float n1[3], d0[3], d1[3];
                cell_normal(&cn[a], n1);
                const float N1 = NORM3(n1);
                if (use_l1) {
                    const float L[3] = {n0[0] / N0 - n1[0] / N1, n0[1] / N0 - n1[1] / N1, n0[2] / N0 - n1[2] / N1};
                    float s[3];
                    for (int k = 0; k < 3; ++k) s[k] = (L[k] > 0.f) ? 1.f : (L[k] == 0.f ? 0.f : -1.f);
                    d0[0] = s[0] * (-SQR(n0[0]) / CUB(N0) + 1.f / N0) + s[1] * (-n0[0] * n0[1] / CUB(N0)) + s[2] * (-n0[0] * n0[2] / CUB(N0));
                    d0[1] = s[0] * (-n0[0] * n0[1] / CUB(N0)) + s[1] * (-SQR(n0[1]) / CUB(N0) + 1.f / N0) + s[2] * (-n0[1] * n0[2] / CUB(N0));
                    d0[2] = s[0] * (-n0[0] * n0[2] / CUB(N0)) + s[1] * (-n0[1] * n0[2] / CUB(N0)) + s[2] * (-SQR(n0[2]) / CUB(N0) + 1.f / N0);
                    d1[0] = s[0] * (SQR(n1[0]) / CUB(N1) - 1.f / N1) + s[1] * (n1[0] * n1[1] / CUB(N1)) + s[2] * (n1[0] * n1[2] / CUB(N1));
                    d1[1] = s[0] * (n1[0] * n1[1] / CUB(N1)) + s[1] * (SQR(n1[1]) / CUB(N1) - 1.f / N1) + s[2] * (n1[1] * n1[2] / CUB(N1));
                    d1[2] = s[0] * (n1[0] * n1[2] / CUB(N1)) + s[1] * (n1[1] * n1[2] / CUB(N1)) + s[2] * (SQR(n1[2]) / CUB(N1) - 1.f / N1);
                } else {
                    pair_l2(n0, N0, n1, N1, d0, d1);
                }
                const float sc = scale * 1.f / norm_count;
                scatter_normal(&c0, d0, sc, mask, grad);
                scatter_normal(&cn[a], d1, sc, mask, grad);
            }
        }
}

/* surf_sign_change_grad_sparse_kernel :895-977 + host :1429-1466, with the loop counter started at 0 (the reference leaves it
 * uninitialised: undefined behaviour) */
void oracle_surf_sign_change_grad_sparse(const int32_t *links, const int32_t *size, const float *data, int n_cols,
                                         const int32_t *cells, int64_t n_cells, uint8_t *mask, int start_dim, int end_dim,
                                         float scale, float *grad) {
    float sc[3];
    ray_scale(size, sc);
    scale = scale / (float)(int)n_cells;
    for (int64_t c = 0; c < n_cells; ++c)
        for (int idx = start_dim; idx < end_dim; ++idx) {
            const int64_t xyz = cells[c];
            const int z = (int)(xyz % size[2]);
            const int64_t xy = xyz / size[2];
            const int y = (int)(xy % size[1]), x = (int)(xy / size[1]);
            const int32_t l000 = links[xyz];
            if (l000 < 0) continue;
            const int32_t ln[3] = {x + 1 < size[0] ? LNK(x + 1, y, z) : -1, y + 1 < size[1] ? LNK(x, y + 1, z) : -1,
                                   z + 1 < size[2] ? LNK(x, y, z + 1) : -1};
            const float v000 = data[(int64_t)l000 * n_cols + idx];
            float g0 = 0.f, gn[3] = {0.f, 0.f, 0.f}, valid = 0.f;
            for (int i = 0; i < 3; ++i) {
                if (ln[i] < 0) continue;
                valid += 1.f;
                const float vi = data[(int64_t)ln[i] * n_cols + idx];
                if (v000 * vi < 0.f) {
                    g0 += ((v000 >= 0.f) ? 1.f : -1.f) * sc[i];
                    gn[i] += ((vi >= 0.f) ? 1.f : -1.f) * sc[i];
                }
            }
            if (valid == 0.f) continue;
            const float a = g0 / valid * scale;
            if (a != 0.f) { grad[(int64_t)l000 * n_cols + idx] += a; if (mask) mask[l000] = 1; }
            for (int i = 0; i < 3; ++i) {
                const float b = gn[i] / valid * scale;
                if (ln[i] >= 0 && b != 0.f) { grad[(int64_t)ln[i] * n_cols + idx] += b; if (mask) mask[ln[i]] = 1; }
            }
        }
}

/* ---- dense surface_normal_grad, loss_kernel.cu:245-396 + host :1289-1325: every cell of the (size - 1)^3 lattice, column idx of
 * a multi-column tensor, always the connectivity test, squared-difference form, no mask; the zero test is on the unscaled
 * corner weight (_add_surface_grad :187-242) ---- */
static int load_cell_col(const int32_t *links, const float *data, int n_cols, int idx, const int32_t *size, int x, int y, int z,
                         Cell8 *c) {
    if (!((x < size[0] - 1) && (y < size[1] - 1) && (z < size[2] - 1))) return 0;
    for (int k = 0; k < 8; ++k) {
        c->l[k] = LNK(x + (k >> 2), y + ((k >> 1) & 1), z + (k & 1));
        if (c->l[k] < 0) return 0;
    }
    for (int k = 0; k < 8; ++k) c->s[k] = data[(int64_t)c->l[k] * n_cols + idx];
    return 1;
}
static void scatter_normal_col(const Cell8 *c, const float *g, float scale, int n_cols, int idx, float *grad) {
    for (int k = 0; k < 8; ++k) {
        const float sx = (k & 4) ? 0.25f : -0.25f, sy = (k & 2) ? 0.25f : -0.25f, sz = (k & 1) ? 0.25f : -0.25f;
        const float gk = sx * g[0] + sy * g[1] + sz * g[2];
        if (gk != 0.f) ATOMIC_ADD(grad[(int64_t)c->l[k] * n_cols + idx], gk * scale);
    }
}
void oracle_surface_normal_grad(const int32_t *links, const int32_t *size, const float *data, int n_cols, float lv, int start_dim,
                                int end_dim, float scale, float *grad) {
    const int nl = (size[0] - 1) * (size[1] - 1) * (size[2] - 1);
    scale = scale / nl;
#pragma omp parallel for schedule(static) if (nl > OMP_MIN_CELLS)
    for (int64_t xyz = 0; xyz < nl; ++xyz)
        for (int idx = start_dim; idx < end_dim; ++idx) {
            const int z = (int)(xyz % (size[2] - 1));
            const int64_t xy = xyz / (size[2] - 1);
            const int y = (int)(xy % (size[1] - 1)), x = (int)(xy / (size[1] - 1));
            Cell8 c0, cn[3];
            if (!load_cell_col(links, data, n_cols, idx, size, x, y, z, &c0)) continue;
            float n0[3];
            cell_normal(&c0, n0);
            int use[3];
            use[2] = load_cell_col(links, data, n_cols, idx, size, x, y, z + 1, &cn[2]) && face_connected(c0.s[1], c0.s[3], c0.s[5], c0.s[7], lv);
            use[1] = load_cell_col(links, data, n_cols, idx, size, x, y + 1, z, &cn[1]) && face_connected(c0.s[2], c0.s[3], c0.s[6], c0.s[7], lv);
            use[0] = load_cell_col(links, data, n_cols, idx, size, x + 1, y, z, &cn[0]) && face_connected(c0.s[4], c0.s[5], c0.s[6], c0.s[7], lv);
            const int norm_count = use[0] + use[1] + use[2];
            const float N0 = NORM3(n0);
            for (int a = 0; a < 3; ++a) {
                if (!use[a]) continue;
                float n1[3], d0[3], d1[3];
                cell_normal(&cn[a], n1);
                pair_l2(n0, N0, n1, NORM3(n1), d0, d1);
                const float sc = scale * 1.f / norm_count;
                scatter_normal_col(&c0, d0, sc, n_cols, idx, grad);
                scatter_normal_col(&cn[a], d1, sc, n_cols, idx, grad);
            }
        }
}

/* ---- lumisphere_tv_grad_sparse, loss_kernel.cu:1067-1177 + host :1661-1697: TV of the radiance seen from ONE direction
 * (basis values sv) plus the change towards a perturbed direction (basis values su), per colour channel; cells are decoded on
 * the (size - 1) lattice; a cell is skipped only where its link is exactly 0 (:1110; a missing corner contributes v000 = 0) ---- */
void oracle_lumisphere_tv_grad_sparse(const int32_t *links, const int32_t *size, const float *sh, int sh_dim, int basis_dim,
                                      const int32_t *cells, int64_t n_cells, const float *sv, const float *su, float scale,
                                      float dir_factor, uint8_t *mask, float *grad) {
    float sc[3];
    ray_scale(size, sc);
    scale = scale / (float)(int)n_cells;
    const int n_col = sh_dim / basis_dim;
    for (int64_t c = 0; c < n_cells; ++c) {
        const int xyz = cells[c];
        const int z = xyz % (size[2] - 1);
        const int xy = xyz / (size[2] - 1);
        const int y = xy % (size[1] - 1), x = xy / (size[1] - 1);
        const int32_t l0 = LNK(x, y, z);
        if (l0 == 0) continue;
        const int32_t ln[3] = {LNK(x + 1, y, z), LNK(x, y + 1, z), LNK(x, y, z + 1)};
        for (int col = 0; col < n_col; ++col) {
            /* per-channel sums in the warp's shuffle-down order clamped to the segment (HeadSegmentedSum) */
            float a0[16], an[3][16], au[16];
            for (int k = 0; k < basis_dim; ++k) {
                const int idx = col * basis_dim + k;
                const float v000 = l0 >= 0 ? sh[(int64_t)l0 * sh_dim + idx] : 0.f;
                a0[k] = v000 * sv[k];
                au[k] = v000 * su[k];
                for (int i = 0; i < 3; ++i) an[i][k] = (ln[i] >= 0 ? sh[(int64_t)ln[i] * sh_dim + idx] : v000) * sv[k];
            }
            for (int off = 1; off < 16; off <<= 1)
                for (int k = 0; k + off < basis_dim; ++k) {
                    a0[k] += a0[k + off];
                    au[k] += au[k + off];
                    for (int i = 0; i < 3; ++i) an[i][k] += an[i][k + off];
                }
            float dx = (an[0][0] - a0[0]) * sc[0], dy = (an[1][0] - a0[0]) * sc[1], dz = (an[2][0] - a0[0]) * sc[2];
            float du = (au[0] - a0[0]) * dir_factor;
            const float idelta = scale * (1.f / sqrtf(1e-9f + dx * dx + dy * dy + dz * dz + du * du));
            dx *= sc[0]; dy *= sc[1]; dz *= sc[2]; du *= dir_factor;
            for (int k = 0; k < basis_dim; ++k) {
                const int idx = col * basis_dim + k;
                const float s = sv[k];
                const float sm = -dx * s - dy * s - dz * s + du * (su[k] - s);
                const float vals[4] = {sm, dx * s, dy * s, dz * s};
                const int32_t ls[4] = {l0, ln[0], ln[1], ln[2]};
                for (int j = 0; j < 4; ++j)
                    if (ls[j] >= 0 && vals[j] != 0.f) {
                        grad[(int64_t)ls[j] * sh_dim + idx] += vals[j] * idelta;
                        if (mask) mask[ls[j]] = 1;
                    }
            }
        }
    }
}
