/* oracle/bigru_ref.c - plain-C CPU restatement of the biGRU hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py.  The product never calls into this file.
 *
 * Restates (float storage, double accumulation):
 *   - model forward        /root/reference/biGRU_model.py:63-138  (eval mode / dropout p=0)
 *     whose recurrent cell is torch.nn.GRU (third-party; /root/reference/requirements.txt:11
 *     pins torch==1.2.0).  Published cell equations, gate row order r|z|n:
 *        r = s(Wir x + bir + Whr h + bhr)      z = s(Wiz x + biz + Whz h + bhz)
 *        n = tanh(Win x + bin + r*(Whn h + bhn))   h' = (1-z)*n + z*h
 *   - its reverse-mode derivative (the reference gets it from autograd, biGRU_model.py:204)
 *   - clip_grad_norm_ + Adam   biGRU_model.py:208-210
 *   - sliding-window collation + min/max normalisation  sql_pytorch_dataloader.py:8-18,239-245
 *   - the three losses the reference's callers plug in (CE for the benchmark configs,
 *     BCEWithLogits notebook raw :1192, MultiLabelSoftMargin predict.py:94)
 *
 * Flat parameter order (shared with the product's C-ABI, include/bigru_b200.h):
 *   for l in [0,L): for d in [0,D): w_ih[3H,I_l] w_hh[3H,H] b_ih[3H] b_hh[3H] ; lin_w[C,3H] lin_b[C]
 *   I_0 = F, I_l = D*H.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int B, T, F, H, L, C, D; } dims_t;

static int64_t in_size(const dims_t* d, int l) { return l == 0 ? d->F : (int64_t)d->D * d->H; }

static int64_t layer_dir_offset(const dims_t* d, int l, int dir) {
    int64_t off = 0;
    for (int ll = 0; ll < d->L; ++ll)
        for (int dd = 0; dd < d->D; ++dd) {
            if (ll == l && dd == dir) return off;
            off += 3LL * d->H * in_size(d, ll) + 3LL * d->H * d->H + 6LL * d->H;
        }
    return off;
}

int64_t bigru_ref_param_count(int F, int H, int L, int C, int D) {
    dims_t d = {0, 0, F, H, L, C, D};
    return layer_dir_offset(&d, L, 0) + 3LL * H * C + C;
}

static double sigm(double a) { return 1.0 / (1.0 + exp(-a)); }

/* Workspace the backward needs; allocated by forward when stash != NULL.
 * stash layout (doubles): per layer: out[B,T,D*H]; per (l,dir): r,z,n,hn,hprev each [B,T,H];
 * then cat[B,3H] and arg[B,H] (as double). */
static int64_t stash_doubles(const dims_t* d) {
    int64_t BT = (int64_t)d->B * d->T;
    return (int64_t)d->L * (BT * d->D * d->H + (int64_t)d->D * 5 * BT * d->H) + (int64_t)d->B * 4 * d->H;
}
int64_t bigru_ref_stash_doubles(int B, int T, int F, int H, int L, int C, int D) {
    dims_t d = {B, T, F, H, L, C, D};
    return stash_doubles(&d);
}

/* forward: x[B,T,F], h0 (nullable) [L*D,B,H] -> logits[B,C], hn_out (nullable) [L*D,B,H] */
int bigru_ref_forward(int B, int T, int F, int H, int L, int C, int D,
                      const float* params, const float* x, const float* h0,
                      float* logits, float* hn_out, double* stash) {
    dims_t dm = {B, T, F, H, L, C, D};
    int64_t BT = (int64_t)B * T;
    int own = 0;
    if (!stash) { stash = (double*)malloc(sizeof(double) * stash_doubles(&dm)); own = 1; if (!stash) return -1; }
    double* sp = stash;
    const double* inp_d = NULL;   /* previous layer's output (double) */
    double* out = NULL;
    for (int l = 0; l < L; ++l) {
        int64_t I = in_size(&dm, l);
        out = sp; sp += BT * D * H;
        for (int dir = 0; dir < D; ++dir) {
            const float* w_ih = params + layer_dir_offset(&dm, l, dir);
            const float* w_hh = w_ih + 3LL * H * I;
            const float* b_ih = w_hh + 3LL * H * H;
            const float* b_hh = b_ih + 3LL * H;
            double *R = sp, *Z = R + BT * H, *N = Z + BT * H, *HN = N + BT * H, *HP = HN + BT * H;
            sp += 5 * BT * H;
            #pragma omp parallel for schedule(static)
            for (int b = 0; b < B; ++b) {
                double* h = (double*)malloc(sizeof(double) * 2 * H);
                double* hnew = h + H;
                for (int j = 0; j < H; ++j) h[j] = h0 ? (double)h0[((int64_t)(l * D + dir) * B + b) * H + j] : 0.0;
                for (int s = 0; s < T; ++s) {
                    int t = dir == 0 ? s : T - 1 - s;
                    int64_t row = (int64_t)b * T + t;
                    for (int j = 0; j < H; ++j) {
                        double gi[3], gh[3];
                        for (int g = 0; g < 3; ++g) {
                            const float* wi = w_ih + (int64_t)(g * H + j) * I;
                            double a = b_ih[g * H + j];
                            if (l == 0) { const float* xr = x + row * F; for (int k = 0; k < I; ++k) a += (double)wi[k] * xr[k]; }
                            else { const double* xr = inp_d + row * I; for (int k = 0; k < I; ++k) a += (double)wi[k] * xr[k]; }
                            gi[g] = a;
                            const float* wh = w_hh + (int64_t)(g * H + j) * H;
                            double c = b_hh[g * H + j];
                            for (int k = 0; k < H; ++k) c += (double)wh[k] * h[k];
                            gh[g] = c;
                        }
                        double r = sigm(gi[0] + gh[0]), z = sigm(gi[1] + gh[1]);
                        double n = tanh(gi[2] + r * gh[2]);
                        hnew[j] = (1.0 - z) * n + z * h[j];
                        R[row * H + j] = r; Z[row * H + j] = z; N[row * H + j] = n; HN[row * H + j] = gh[2];
                        HP[row * H + j] = h[j];
                    }
                    for (int j = 0; j < H; ++j) { h[j] = hnew[j]; out[row * D * H + dir * H + j] = hnew[j]; }
                }
                if (hn_out) for (int j = 0; j < H; ++j) hn_out[((int64_t)(l * D + dir) * B + b) * H + j] = (float)h[j];
                free(h);
            }
        }
        inp_d = out;
    }
    /* head: biGRU_model.py:111-137 */
    double* cat = sp; double* arg = cat + (int64_t)B * 3 * H;
    const float* lin_w = params + layer_dir_offset(&dm, L, 0);
    const float* lin_b = lin_w + 3LL * H * C;
    for (int b = 0; b < B; ++b) {
        for (int j = 0; j < H; ++j) {
            double last = out[((int64_t)b * T + (T - 1)) * D * H + j];
            if (D == 2) last += out[((int64_t)b * T + 0) * D * H + H + j];
            double mx = -INFINITY, sum = 0.0; int am = 0;
            for (int t = 0; t < T; ++t) {
                double s = out[((int64_t)b * T + t) * D * H + j];
                if (D == 2) s += out[((int64_t)b * T + t) * D * H + H + j];
                if (s > mx) { mx = s; am = t; }          /* strict >: first maximum wins */
                sum += s;
            }
            cat[(int64_t)b * 3 * H + j] = last;
            cat[(int64_t)b * 3 * H + H + j] = mx;
            cat[(int64_t)b * 3 * H + 2 * H + j] = sum / (double)T;
            arg[(int64_t)b * H + j] = (double)am;
        }
        for (int c = 0; c < C; ++c) {
            double a = lin_b[c];
            for (int k = 0; k < 3 * H; ++k) a += (double)lin_w[(int64_t)c * 3 * H + k] * cat[(int64_t)b * 3 * H + k];
            logits[(int64_t)b * C + c] = (float)a;
        }
    }
    if (own) free(stash);
    return 0;
}

/* backward: needs the stash written by bigru_ref_forward.  grads: flat, same order as params.
 * dx nullable [B,T,F]; dh0 nullable [L*D,B,H]. */
int bigru_ref_backward(int B, int T, int F, int H, int L, int C, int D,
                       const float* params, const float* x, const double* stash,
                       const float* dlogits, float* grads, float* dx, float* dh0) {
    dims_t dm = {B, T, F, H, L, C, D};
    int64_t BT = (int64_t)B * T;
    int64_t P = bigru_ref_param_count(F, H, L, C, D);
    double* g = (double*)calloc(P, sizeof(double));
    /* locate stash pieces */
    const double** outs = (const double**)malloc(sizeof(double*) * L);
    const double** gates = (const double**)malloc(sizeof(double*) * L * D);
    const double* sp = stash;
    for (int l = 0; l < L; ++l) {
        outs[l] = sp; sp += BT * D * H;
        for (int dir = 0; dir < D; ++dir) { gates[l * D + dir] = sp; sp += 5 * BT * H; }
    }
    const double* cat = sp; const double* arg = cat + (int64_t)B * 3 * H;
    int64_t lin_off = layer_dir_offset(&dm, L, 0);
    const float* lin_w = params + lin_off;
    double* dcat = (double*)calloc((size_t)B * 3 * H, sizeof(double));
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            double dl = dlogits[(int64_t)b * C + c];
            g[lin_off + 3LL * H * C + c] += dl;
            for (int k = 0; k < 3 * H; ++k) {
                g[lin_off + (int64_t)c * 3 * H + k] += dl * cat[(int64_t)b * 3 * H + k];
                dcat[(int64_t)b * 3 * H + k] += dl * lin_w[(int64_t)c * 3 * H + k];
            }
        }
    /* dout for the top layer */
    double* dout = (double*)calloc((size_t)BT * D * H, sizeof(double));
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < H; ++j) {
            double dmax = dcat[(int64_t)b * 3 * H + H + j], davg = dcat[(int64_t)b * 3 * H + 2 * H + j] / (double)T;
            int am = (int)arg[(int64_t)b * H + j];
            for (int t = 0; t < T; ++t) {
                double v = davg + (t == am ? dmax : 0.0);
                for (int dir = 0; dir < D; ++dir) dout[((int64_t)b * T + t) * D * H + dir * H + j] = v;
            }
        }
    for (int l = L - 1; l >= 0; --l) {
        int64_t I = in_size(&dm, l);
        double* dinp = (double*)calloc((size_t)BT * I, sizeof(double));
        for (int dir = 0; dir < D; ++dir) {
            int64_t off = layer_dir_offset(&dm, l, dir);
            const float* w_ih = params + off;
            const float* w_hh = w_ih + 3LL * H * I;
            int64_t o_wih = off, o_whh = off + 3LL * H * I, o_bih = o_whh + 3LL * H * H, o_bhh = o_bih + 3LL * H;
            const double *R = gates[l * D + dir], *Z = R + BT * H, *N = Z + BT * H, *HN = N + BT * H, *HP = HN + BT * H;
            double* dgi = (double*)malloc(sizeof(double) * BT * 3 * H);
            double* dgh = (double*)malloc(sizeof(double) * BT * 3 * H);
            #pragma omp parallel for schedule(static)
            for (int b = 0; b < B; ++b) {
                double* dh = (double*)calloc(2 * (size_t)H, sizeof(double));
                double* dhn = dh + H;
                if (l == L - 1) for (int j = 0; j < H; ++j) dh[j] = dcat[(int64_t)b * 3 * H + j];
                for (int s = 0; s < T; ++s) {
                    int t = dir == 0 ? T - 1 - s : s;
                    int64_t row = (int64_t)b * T + t;
                    for (int j = 0; j < H; ++j) {
                        double d = dh[j] + dout[row * D * H + dir * H + j];
                        double r = R[row * H + j], z = Z[row * H + j], n = N[row * H + j];
                        double hn = HN[row * H + j], hp = HP[row * H + j];
                        double dan = d * (1.0 - z) * (1.0 - n * n);
                        double dar = dan * hn * r * (1.0 - r);
                        double daz = d * (hp - n) * z * (1.0 - z);
                        dgi[row * 3 * H + j] = dar; dgi[row * 3 * H + H + j] = daz; dgi[row * 3 * H + 2 * H + j] = dan;
                        dgh[row * 3 * H + j] = dar; dgh[row * 3 * H + H + j] = daz; dgh[row * 3 * H + 2 * H + j] = dan * r;
                        dhn[j] = d * z;
                    }
                    for (int k = 0; k < H; ++k) {
                        double a = dhn[k];
                        for (int q = 0; q < 3 * H; ++q) a += dgh[row * 3 * H + q] * w_hh[(int64_t)q * H + k];
                        dh[k] = a;
                    }
                }
                if (dh0) for (int j = 0; j < H; ++j) dh0[((int64_t)(l * D + dir) * B + b) * H + j] = (float)dh[j];
                free(dh);
            }
            /* weight / bias / input gradients */
            #pragma omp parallel for schedule(static)
            for (int q = 0; q < 3 * H; ++q) {
                double sb_i = 0.0, sb_h = 0.0;
                for (int64_t row = 0; row < BT; ++row) {
                    double a = dgi[row * 3 * H + q], c = dgh[row * 3 * H + q];
                    sb_i += a; sb_h += c;
                    if (l == 0) { const float* xr = x + row * F; for (int k = 0; k < I; ++k) g[o_wih + (int64_t)q * I + k] += a * xr[k]; }
                    else { const double* xr = outs[l - 1] + row * I; for (int k = 0; k < I; ++k) g[o_wih + (int64_t)q * I + k] += a * xr[k]; }
                    const double* hp = HP + row * H;
                    for (int k = 0; k < H; ++k) g[o_whh + (int64_t)q * H + k] += c * hp[k];
                }
                g[o_bih + q] = sb_i; g[o_bhh + q] = sb_h;
            }
            #pragma omp parallel for schedule(static)
            for (int64_t row = 0; row < BT; ++row)
                for (int q = 0; q < 3 * H; ++q) {
                    double a = dgi[row * 3 * H + q];
                    const float* wi = w_ih + (int64_t)q * I;
                    for (int k = 0; k < I; ++k) dinp[row * I + k] += a * wi[k];
                }
            free(dgi); free(dgh);
        }
        free(dout);
        dout = dinp;
    }
    if (dx) for (int64_t i = 0; i < BT * F; ++i) dx[i] = (float)dout[i];
    free(dout);
    for (int64_t i = 0; i < P; ++i) grads[i] = (float)g[i];
    free(g); free(dcat); free(outs); free(gates);
    return 0;
}

/* ---- losses: value (mean reduction) + dlogits --------------------------------------- */
double bigru_ref_loss_ce(int B, int C, const float* logits, const int64_t* target, float* dlogits, double scale) {
    double loss = 0.0;
    for (int b = 0; b < B; ++b) {
        const float* lg = logits + (int64_t)b * C;
        double m = lg[0]; for (int c = 1; c < C; ++c) if (lg[c] > m) m = lg[c];
        double s = 0.0; for (int c = 0; c < C; ++c) s += exp(lg[c] - m);
        double lse = m + log(s);
        loss += lse - lg[target[b]];
        if (dlogits) for (int c = 0; c < C; ++c)
            dlogits[(int64_t)b * C + c] = (float)((exp(lg[c] - lse) - (c == target[b] ? 1.0 : 0.0)) * scale);
    }
    return loss * scale;
}

/* torch.nn.BCEWithLogitsLoss(weight[C] nullable, pos_weight[C] nullable), mean over B*C */
double bigru_ref_loss_bce(int B, int C, const float* logits, const float* target, const float* weight,
                          const float* pos_weight, float* dlogits, double scale) {
    double loss = 0.0;
    for (int b = 0; b < B; ++b) for (int c = 0; c < C; ++c) {
        double x = logits[(int64_t)b * C + c], y = target[(int64_t)b * C + c];
        double pw = pos_weight ? pos_weight[c] : 1.0, w = weight ? weight[c] : 1.0;
        double sp_neg = (x > 0 ? 0.0 : -x) + log1p(exp(-fabs(x)));   /* softplus(-x) = -log sigmoid(x) */
        double lw = 1.0 + (pw - 1.0) * y;
        loss += w * ((1.0 - y) * x + lw * sp_neg);
        if (dlogits) {
            double s = sigm(x);
            dlogits[(int64_t)b * C + c] = (float)(w * ((1.0 - y) - lw * (1.0 - s)) * scale);
        }
    }
    return loss * scale;
}

/* torch.nn.MultiLabelSoftMarginLoss(): mean over classes, then mean over batch */
double bigru_ref_loss_mlsm(int B, int C, const float* logits, const float* target, float* dlogits, double scale) {
    return bigru_ref_loss_bce(B, C, logits, target, NULL, NULL, dlogits, scale);
}

/* ---- clip_grad_norm_ + Adam (biGRU_model.py:208-210; torch.optim.Adam defaults) ------- */
double bigru_ref_clip_adam(int64_t P, float* params, float* grads, float* m, float* v, double clip,
                           double lr, double b1, double b2, double eps, int step) {
    double ss = 0.0;
    for (int64_t i = 0; i < P; ++i) ss += (double)grads[i] * grads[i];
    double norm = sqrt(ss);
    double coef = clip / (norm + 1e-6);
    if (coef > 1.0) coef = 1.0;
    double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    for (int64_t i = 0; i < P; ++i) {
        float gq = (float)(grads[i] * (float)coef);
        grads[i] = gq;
        m[i] = (float)(b1 * m[i] + (1.0 - b1) * gq);
        v[i] = (float)(b2 * v[i] + (1.0 - b2) * (double)gq * gq);
        double denom = sqrt((double)v[i]) / sqrt(bc2) + eps;
        params[i] = (float)(params[i] - (lr / bc1) * (m[i] / denom));
    }
    return norm;
}

/* ---- window collation: out[b,t,:] = (src[start+b+t,:]-min)/(max-min)
 * sql_pytorch_dataloader.py:239 (normalise) and :8-18,:243-245 (stride-1 windows) */
void bigru_ref_window_gather_norm(const float* src, const float* xmin, const float* xmax,
                                  int64_t start, int B, int T, int F, float* out) {
    for (int b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t)
            for (int f = 0; f < F; ++f) {
                float v = src[(start + b + t) * F + f];
                out[((int64_t)b * T + t) * F + f] = (v - xmin[f]) / (xmax[f] - xmin[f]);
            }
}
