/* medseg_b200.h — C ABI of libmedseg_b200.so (hand-written sm_100a CUDA for the MedicalSeg VNet hot path).
 *
 * The reference (PaddleCV-SIG/MedicalSeg) has NO FFI / operator-plugin interface: every device op it runs is a
 * PaddlePaddle kernel reached through paddle.nn.  The entry points below are therefore the
 * operators that medicalseg/models/vnet.py, medicalseg/models/losses/*.py, cvlibs/config.py (Momentum) and
 * tools/preprocess_utils/*.py reach inside Paddle / NumPy / CuPy; each one cites the reference call site it
 * replaces (paths relative to /root/reference).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative msb_status; the message is read with
 *    msb_last_error_string() (thread-local).  Nothing throws across the ABI, nothing calls exit().
 *  - the caller owns every buffer (PyTorch caching allocator); the library allocates nothing persistent.
 *  - all work is enqueued on the cudaStream_t passed as `void* stream`; no function synchronises.
 *  - activations use the blocked-8 layout "B8": [N][C/8][D][H][W][8] (8 channels innermost), element type
 *    bf16 or f32.  A msb_tensor may be a channel-slice view of a wider buffer (concat without copy).
 *  - boundary tensors (network input, logits, labels, parameters, gradients) keep the reference layouts:
 *    NCDHW f32 / [N,D,H,W] int32 / Paddle parameter shapes.
 */
#ifndef MEDSEG_B200_H_
#define MEDSEG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSB_VERSION 100

typedef enum {
  MSB_OK = 0,
  MSB_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  MSB_ERR_CUDA = -2,      /* CUDA runtime or driver error */
  MSB_ERR_UNSUPPORTED = -3
} msb_status;

typedef enum { MSB_F32 = 0, MSB_BF16 = 1 } msb_dtype;

/* View of a B8 activation: `ptr` addresses (n=0, first 8-channel plane of the view). */
typedef struct {
  void* ptr;
  int64_t n_stride; /* elements between consecutive n in the underlying buffer (= C8_total*D*H*W*8) */
  int32_t c;        /* channels of this view, multiple of 8 */
  int32_t dtype;    /* msb_dtype */
} msb_tensor;

typedef struct { int32_t d, h, w; } msb_dim3;

int msb_version(void);
const char* msb_last_error_string(void);

/* ---- layout converters (boundary) ------------------------------------------------------------------ */
/* Tile scheduling of the persistent tensor-core kernels (5x5x5 fwd / dgrad, both weight-gradient kernels, the strided
 * convs): 0 = static round-robin over the grid (default), 1 = CTAs fetch tiles from a self-resetting atomic counter in
 * device memory - an experiment for the data-parallel step (medicalseg/core/train.py:81-88), where NCCL's all-reduce
 * CTAs occupy SMs while backward runs.  Measured to gain nothing on 2 and 8 GPUs, so the default build compiles it out
 * (msb_set_tile_scheduler(1) -> MSB_ERR_UNSUPPORTED); `MSB_DYNAMIC_TILES=1 python -m medicalseg_b200.build --force`
 * builds it in. */
int msb_set_tile_scheduler(int dynamic);
/* Clears a device buffer on `stream` (cudaMemsetAsync: a memset node inside a CUDA-graph capture).  Replaces
 * `Layer.clear_gradients()` / `paddle.zeros` on the train path (medicalseg/core/train.py:155). */
int msb_zero(void* ptr, size_t bytes, void* stream);
/* NCDHW f32 [N,C,S] -> B8 view (channels >= C zero-filled up to dst.c).  vnet.py:256 (network input side). */
int msb_to_blocked(const float* src, int n, int c, int64_t s, msb_tensor dst, void* stream);
/* B8 view -> NCDHW f32 [N,C,S]. */
int msb_from_blocked(msb_tensor src, float* dst, int n, int c, int64_t s, void* stream);

/* ---- BatchNorm3D + PReLU (+ tile-add / residual-add + second PReLU) --------------------------------
 * Replaces nn.BatchNorm3D / nn.PReLU / paddle.add / Tensor.tile at vnet.py:41,74-79,107-111,149-154,173.
 * groups = 1 -> batch statistics (reference semantics); groups = N -> per-instance statistics.       */
/* sums[2][groups][C] (double) += per-channel sum and sum of squares of x. */
int msb_bn_stats(msb_tensor x, int n, int64_t s, int groups, double* sums, void* stream);
/* bnbuf[4][groups][C] f32 = scale, shift, mean, invstd.  training: from sums (biased variance), running stats
 * updated as running = momentum*running + (1-momentum)*batch (Paddle convention); eval: from running stats. */
int msb_bn_finalize(const double* sums, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, int training,
                    int c, int groups, float* bnbuf, void* stream);
/* out = act2( act1( bn(y) [+ tile(tile_src)] ) [+ residual] ); act = PReLU(alpha).  residual.ptr==NULL and
 * alpha2==NULL skip the second stage; tile_src (NCDHW f32 [N,tile_c,S], channel c reads c % tile_c —
 * Tensor.tile, vnet.py:76-78) may be NULL. */
int msb_bn_act_fwd(msb_tensor y, msb_tensor out, msb_tensor residual, const float* tile_src, int tile_c,
                   const float* bnbuf, const float* alpha1, const float* alpha2,
                   int n, int64_t s, int groups, void* stream);
/* msb_bn_finalize + msb_bn_act_fwd in one launch (same arguments; bnbuf is written for the backward kernels and the
 * running statistics are updated by one designated block). */
int msb_bn_fwd_fused(msb_tensor y, msb_tensor out, msb_tensor residual, const float* tile_src, int tile_c,
                     const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int training, float* bnbuf, const float* alpha1,
                     const float* alpha2, int n, int64_t s, int groups, void* stream);
/* red[4][groups][C] (double) += sum g1, sum g1*xhat, dalpha1, dalpha2 (g1 = grad at the BN output). */
int msb_bn_act_bwd_reduce(msb_tensor y, msb_tensor residual, const float* tile_src, int tile_c, msb_tensor gout,
                          const float* bnbuf, const float* alpha1, const float* alpha2,
                          int n, int64_t s, int groups, double* red, void* stream);
/* dy = BN input gradient; dres (+)= gradient of the residual branch (dres.ptr may be NULL);
 * dgamma/dbeta/dalpha1/dalpha2 (f32, += when accumulate_params) from red. */
int msb_bn_act_bwd_apply(msb_tensor y, msb_tensor residual, const float* tile_src, int tile_c, msb_tensor gout,
                         const float* bnbuf, const float* alpha1, const float* alpha2,
                         const double* red, double count, int training, msb_tensor dy, msb_tensor dres,
                         int dres_accumulate, float* dgamma, float* dbeta, float* dalpha1, float* dalpha2,
                         int n, int64_t s, int groups, void* stream);

/* dst = (accumulate ? dst : 0) + scale[n][c] * src.  nn.Dropout3D fwd/bwd (vnet.py:108,149-150) with the
 * explicit [N,C] mask (0 or 2); scale==NULL means 1 (plain copy / add). */
int msb_channel_scale(msb_tensor src, msb_tensor dst, const float* scale, int n, int64_t s, int accumulate,
                      void* stream);

/* ---- OutputTransition.conv2 (1x1x1, vnet.py:169,174) ------------------------------------------------ */
/* logits NCDHW f32 [N,Co,S] = W[Co][Ci] * a + b */
int msb_conv1x1_fwd(msb_tensor a, const float* w, const float* b, float* logits, int n, int ci, int co,
                    int64_t s, void* stream);
/* da = W^T dlogits;  dw += a (x) dlogits;  db += sum dlogits   (dw, db f32, accumulated atomically) */
int msb_conv1x1_bwd(msb_tensor a, const float* w, const float* dlogits, msb_tensor da, float* dw, float* db,
                    int n, int ci, int co, int64_t s, void* stream);

/* ---- strided Conv3D / Conv3DTranspose without padding on tcgen05 tensor cores (general kernel / stride) ------
 * vnet.py:98-99 (down_conv), :133-137 (up_conv) incl. the anisotropic MRISpineSeg kernels (2,2,4)/(2,2,1) and
 * (2,2,2)/(2,2,1) whose windows overlap along the last axis.  mode 0 = gather (conv forward / transposed-conv input
 * gradient; weight [c_out][c_red][kd*kh*kw]), mode 1 = scatter (transposed-conv forward / conv input gradient; weight
 * [c_red][c_out][kd*kh*kw]; needs stride == kernel along d and h, and stride == kernel or stride == 1 along w).
 * big_dims = extents of the LARGE grid; (big - kernel) % stride == 0 per axis.  MSB_ERR_UNSUPPORTED otherwise (callers
 * fall back to msb_conv_strided_*).  msb_conv_tc_packed_bytes returns 0 for unsupported geometry. */
size_t msb_conv_tc_packed_bytes(int c_red_pad, int c_out_pad, msb_dim3 kernel, msb_dim3 stride, int mode);
int msb_conv_tc_pack(const float* w, void* packed, int c_red, int c_out, int mode, int c_red_pad, int c_out_pad,
                     msb_dim3 kernel, msb_dim3 stride, void* stream);
int msb_conv_tc_gather(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                       msb_dim3 big_dims, msb_dim3 kernel, msb_dim3 stride, int groups, double* sums, void* stream);
int msb_conv_tc_scatter(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                        msb_dim3 big_dims, msb_dim3 kernel, msb_dim3 stride, int accumulate, int groups, double* sums,
                        void* stream);
/* dw [small.c][big.c][taps] f32 += weight gradient (both Conv3D [out][in][k] with big = input, and Conv3DTranspose
 * [in][out][k] with big = output gradient); taps * big.c must be a multiple of 128, big.c in {16,32,64,128} */
size_t msb_conv_tc_wgrad_workspace_bytes(int c_big, int c_small, msb_dim3 kernel);
int msb_conv_tc_wgrad(msb_tensor big, msb_tensor small, float* dw, float* dbias, int n, msb_dim3 big_dims,
                      msb_dim3 kernel, msb_dim3 stride, int bias_from_big, void* workspace, size_t workspace_bytes,
                      void* stream);
/* ---- direct (CUDA-core) Conv3D / Conv3DTranspose for the HBM-bound layers ---------------------------
 * vnet.py:67-68 (in_tr 1->16, k5), :98-99 (down_conv k=kernel,s=stride), :133-137 (up_conv, transposed).
 * Weights/grads stay in the reference (Paddle) layouts: conv [Cout,Cin,kD,kH,kW]; convT [Cin,Cout,kD,kH,kW]. */
/* in_tr: x NCDHW f32 [N,1,D,H,W], w [16][1][5][5][5]; out B8 (16 ch); optional BN partial sums (double [2][G][16]) */
int msb_conv_in_fwd(const float* x, const float* w, const float* bias, msb_tensor out, int n, msb_dim3 dims,
                    int groups, double* sums, void* stream);
int msb_conv_in_wgrad(const float* x, msb_tensor dy, float* dw, float* dbias, int n, msb_dim3 dims, void* stream);
/* c_*_real: channel counts of the weight tensor when the views are zero-padded wider (0 = same as the view).
 * gather form (zero padding `pad`): out[o, oc] = bias[oc] + sum_{tap, rc} x[o*s + tap - pad, rc] * w[oc][rc][tap].
 * = forward of nn.Conv3D (w = [Cout,Cin,k]) and input-gradient of nn.Conv3DTranspose (w = [Cin_T,Cout_T,k]). */
int msb_conv_strided_fwd(msb_tensor x, const float* w, const float* bias, msb_tensor out, int n,
                         msb_dim3 in_dims, msb_dim3 kernel, msb_dim3 stride, msb_dim3 pad, int c_red_real,
                         int c_out_real, int groups, double* sums, void* stream);
/* scatter form: out[i, oc] (+)= bias[oc] + sum_{tap, o: o*s+tap-pad==i} sum_rc x[o, rc] * w[rc][oc][tap].
 * = input-gradient of nn.Conv3D (w = [Cout,Cin,k]) and forward of nn.Conv3DTranspose (w = [Cin_T,Cout_T,k]);
 * out_dims = (in-1)*s + k - 2*pad.  Optional BN partial sums of the (rounded) outputs. */
int msb_conv_strided_bwd_data(msb_tensor x, const float* w, const float* bias, msb_tensor out, int n,
                              msb_dim3 out_dims, msb_dim3 kernel, msb_dim3 stride, msb_dim3 pad, int c_red_real,
                              int c_out_real, int accumulate, int groups, double* sums, void* stream);
/* dw[sc][bc][tap] += sum_o big[o*s+tap-pad, bc] * small[o, sc];  dbias += sum small (bias_from_big=0, nn.Conv3D:
 * big = x, small = dy) or sum big (bias_from_big=1, nn.Conv3DTranspose: big = dy, small = x). */
int msb_conv_strided_wgrad(msb_tensor big, msb_tensor small, float* dw, float* dbias, int n,
                           msb_dim3 big_dims, msb_dim3 kernel, msb_dim3 stride, msb_dim3 pad, int c_big_real,
                           int c_small_real, int bias_from_big, void* stream);

/* tensor-core weight gradient for the non-overlapping kernel = stride = (2,2,2) case (bf16 views, even extents) - the
 * wrapper of msb_conv_tc_wgrad: a pointwise tcgen05 weight-gradient GEMM whose M rows are (tap, channel); every tap's
 * sub-lattice of `big` is fetched by a strided TMA box (no space-to-depth copy).  Same result layout as
 * msb_conv_strided_wgrad.  workspace: msb_conv_k2s2_wgrad_workspace_bytes(...) bytes of device scratch. */
size_t msb_conv_k2s2_wgrad_workspace_bytes(int n, int c_big, int c_small, msb_dim3 big_dims);
int msb_conv_k2s2_wgrad(msb_tensor big, msb_tensor small, float* dw, float* dbias, int n, msb_dim3 big_dims,
                        int bias_from_big, void* workspace, size_t workspace_bytes, void* stream);

/* tensor-core forward / input-gradient of the same non-overlapping case (bf16 views, even extents):
 *   gather : out(small grid)[v,co] = bias + sum_{tap,cr} x(big)[2v+tap,cr] * w[co][cr][tap]   (nn.Conv3D forward,
 *            vnet.py:98-99; also the input gradient of nn.Conv3DTranspose with its [Cin][Cout][8] weight)
 *   scatter: out(big grid)[2v+tap,co] (+)= bias + sum_cr x(small)[v,cr] * w[cr][co][tap]      (nn.Conv3DTranspose
 *            forward, vnet.py:133-137; also the input gradient of nn.Conv3D)
 * packed = msb_conv_k2s2_pack(w, ..., mode 0 gather | 1 scatter, c_red_pad = x.c, c_out_pad = pad16(out.c)).
 * Optional BN partial sums (double [2][groups][out.c]) of the rounded outputs, as msb_conv_strided_fwd. */
size_t msb_conv_k2s2_packed_bytes(int c_red_pad, int c_out_pad);
int msb_conv_k2s2_pack(const float* w, void* packed, int c_red, int c_out, int mode, int c_red_pad, int c_out_pad,
                       void* stream);
int msb_conv_k2s2_gather(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                         msb_dim3 big_dims, int groups, double* sums, void* stream);
int msb_conv_k2s2_scatter(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                          msb_dim3 big_dims, int accumulate, int groups, double* sums, void* stream);

/* ---- 5x5x5 Conv3D (pad 2, stride 1) on tcgen05 tensor cores ----------------------------------------
 * vnet.py:36 (LUConv.conv1, 14x), :165-166 (out_tr.conv1).  bf16 operands, f32 accumulation in TMEM.     */
/* packs the Paddle-layout f32 weight [Cout][Cin][125] into the bf16 UMMA operand image.
 * mode 0: forward operand; mode 1: input-gradient operand (taps flipped, Cin/Cout swapped).
 * cin_pad/cout_pad (multiples of 16) are the channel counts of the conv the packed operand will be used in. */
size_t msb_conv_k5_packed_bytes(int cin_pad, int cout_pad);
/* N (output-channel) padding the tensor-core kernel uses for an output view of `cout_view` channels */
int msb_conv_k5_out_pad(int cout_view);
int msb_conv_k5_pack(const float* w, void* packed, int cout, int cin, int mode, int cin_pad, int cout_pad,
                     void* stream);
/* out = conv(x) + bias (bias may be NULL, `cout` real output channels <= out.c); x.c must be a multiple of 16
 * and the packed operand built with cin_pad = x.c, cout_pad = msb_conv_k5_out_pad(out.c).
 * Optional BN partial sums (double [2][groups][out.c]) of the rounded outputs.
 * accumulate: out += result * ch_scale[n][c] (ch_scale may be NULL) - used for input gradients. */
int msb_conv_k5_fwd(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                    msb_dim3 dims, int accumulate, const float* ch_scale, int groups, double* sums, void* stream);
/* Same op with a split-K path for SMALL volumes (fewer 128-voxel tiles than SMs; the 8^3 / 16^3 levels of VNet):
 * the reduction over the input channels is split across CTAs so that every SM streams only a slice of the 4-16 MB
 * weight set; every K slice stores its partial f32 tile into its OWN copy inside `workspace` ([slices][n][c8][S][8])
 * and a finalize kernel adds the copies in slice order - bit-reproducible, no atomics - and applies bias / accumulate
 * / bf16 rounding / BN sums.  msb_conv_k5_fwd_workspace_bytes() returns 0 when the shape uses the regular path
 * (workspace may then be NULL).  The workspace is pure scratch: no initial contents required, contents undefined
 * afterwards, one buffer may serve every layer of a model (calls are stream-ordered). */
size_t msb_conv_k5_fwd_workspace_bytes(int n, int cout_view, msb_dim3 dims, int cin_view);
int msb_conv_k5_fwd_ws(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                       msb_dim3 dims, int accumulate, const float* ch_scale, int groups, double* sums, void* workspace,
                       size_t workspace_bytes, void* stream);
/* Evaluation-mode LUConv in one kernel (vnet.py:41 `relu1(bn1(conv1(x)))`, and :110 / :154 where the block's residual
 * is added before the last PReLU, with BatchNorm using its running statistics - core/val.py:93, model.eval()):
 *   t = prelu((conv(x) + bias) * scale[c] + shift[c], alpha[c]);  out = residual ? prelu(t + residual, alpha2[c]) : t
 * scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale (f32 [>= out.c], like alpha/alpha2);
 * residual: NULL or a bf16 B8 view shaped like out (then alpha2 is required).  bf16 output only.  workspace as for
 * msb_conv_k5_fwd_ws (NULL = regular path). */
int msb_conv_k5_fwd_act(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                        msb_dim3 dims, const float* scale, const float* shift, const float* alpha,
                        const msb_tensor* residual, const float* alpha2, void* workspace, size_t workspace_bytes,
                        void* stream);
/* dw [cout][cin][125] f32 += sum_v x[v+tap] (x) dy[v];  dbias[cout] += sum_v dy[v] (dbias may be NULL).
 * workspace: msb_conv_k5_wgrad_workspace_bytes(cin, cout) bytes of device scratch. */
size_t msb_conv_k5_wgrad_workspace_bytes(int cin, int cout);
int msb_conv_k5_wgrad(msb_tensor x, msb_tensor dy, float* dw, float* dbias, int cout, int cin, int n,
                      msb_dim3 dims, void* workspace, size_t workspace_bytes, void* stream);
/* TAP-MAJOR variants: the f32 master weight / gradient of a 5x5x5 conv stored as [125 taps][cout][cin] - the layout
 * the weight-gradient kernels accumulate in.  msb_conv_k5_wgrad_tm adds straight into dw_tm (no workspace, no memset,
 * no transposition kernel); msb_conv_k5_pack_tm builds the same bf16 operand image as msb_conv_k5_pack from it.  The
 * host layer keeps parameters, gradients and momentum of these layers in this layout and converts only at the
 * state-dict boundary. */
int msb_conv_k5_pack_tm(const float* w_tm, void* packed, int cout, int cin, int mode, int cin_pad, int cout_pad,
                        void* stream);
/* Both operand images of a layer in one pass over the master weight (what the train step calls after every optimizer
 * step: medicalseg/core/train.py:152 `optimizer.step()` changes every weight): packed_f = the mode-0 image built with
 * (f_cin_pad, f_cout_pad), packed_b = the mode-1 image built with (b_cin_pad, b_cout_pad); lo_part = 1 packs w - bf16(w)
 * (3 x bf16 path).  Bit-identical to two msb_conv_k5_pack_tm calls. */
int msb_conv_k5_pack_tm_pair(const float* w_tm, void* packed_f, void* packed_b, int cout, int cin, int lo_part,
                             int f_cin_pad, int f_cout_pad, int b_cin_pad, int b_cout_pad, void* stream);
/* ---- 3 x bf16 fp32 path (BASELINE configs[2]: fp32 storage, tensor-core convolutions) -------------------------
 * An f32 operand a is split into bf16 hi = bf16(a) and lo = bf16(a - hi); conv(x, w) ~ conv(x_hi, w_hi) + conv(x_lo, w_hi)
 * + conv(x_hi, w_lo) with f32 accumulation (relative error ~2^-16; the lo*lo term is dropped).  msb_split_hi_lo splits
 * an f32 B8 activation; msb_conv_k5_pack_tm with mode | 2 packs the lo part of the weights; msb_conv_k5_fwd with an f32
 * B8 `out` view stores f32 and honours `accumulate` / `ch_scale` / `sums` (pass bias on the first and sums on the
 * last of the three passes); msb_conv_k5_wgrad_tm is simply called for the three operand pairs. */
int msb_split_hi_lo(msb_tensor x, msb_tensor hi, msb_tensor lo, int n, int64_t s, void* stream);
int msb_conv_k5_wgrad_tm(msb_tensor x, msb_tensor dy, float* dw_tm, float* dbias, int cout, int cin, int n,
                         msb_dim3 dims, void* stream);
/* ---- w-folded 5x5x1 variant of the 5x5x5 conv for layers with <= 3 real channels on one side -------------
 * vnet.py:67-68 (in_tr.conv1, 1 -> 16: fold_side 0 = input folded) and vnet.py:165-166 (out_tr.conv1, 32 -> classes:
 * fold_side 1 = output folded).  The five kw taps are folded into the zero padding of the 16-channel block:
 *   fold  : F_s(t)[v,(j,c)] = t[v + s*(j-2) e_w, c]            (j = 0..4; 16-channel B8 bf16 result)
 *   unfold: y[v,c] = bias[c] + sum_j P[v + (j-2) e_w,(j,c)]    (+ BN partial sums of the rounded y)
 *   fold_side 0:  conv5(x)  = conv551(F_+1(x));   dW via conv551_wgrad(F_+1(x), dy)
 *   fold_side 1:  conv5(x)  = unfold(conv551(x)); dx = conv551(F_-1(dy)) with the mode-1 operand;
 *                 dW via conv551_wgrad(x, F_-1(dy))
 * so one tcgen05.mma covers all five kw taps (5x fewer MMAs than the plain kernel on these operand-fetch-bound layers). */
int msb_fold_w_f32(const float* x /* NCDHW f32 [n][c_real][D][H][W] */, int c_real, msb_tensor out, int n,
                   msb_dim3 dims, int sign, void* stream);
int msb_fold_w(msb_tensor x, int c_real, msb_tensor out, int n, msb_dim3 dims, int sign, void* stream);
/* p: 16-channel B8 view (f32 or bf16); out: bf16 B8 view whose channels >= c_real are written as zeros;
 * sums: optional BN partial sums double [2][groups][out.c] */
int msb_unfold_w(msb_tensor p, const float* bias, int c_real, msb_tensor out, int n, msb_dim3 dims, int groups,
                 double* sums, void* stream);
/* packs the 5-D Paddle weight [cout][cin][5][5][5] into the folded 5x5x1 operand image (mode as msb_conv_k5_pack) */
size_t msb_conv_k551_packed_bytes(int cin_pad, int cout_pad);
int msb_conv_k551_pack(const float* w, void* packed, int cout, int cin, int mode, int fold_side, int cin_pad,
                       int cout_pad, void* stream);
/* same contract as msb_conv_k5_fwd for the 5x5x1 (kd,kh) kernel, pad (2,2,0); `out` may also be an f32 B8 view
 * (then accumulate = 0 and sums = NULL) */
int msb_conv_k551_fwd(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                      msb_dim3 dims, int accumulate, const float* ch_scale, int groups, double* sums, void* stream);
/* dw [cout][cin][125] f32 (the 5-D weight, cout/cin REAL counts) += weight gradient computed on the folded views */
size_t msb_conv_k551_wgrad_workspace_bytes(int cin, int cout, int fold_side);
int msb_conv_k551_wgrad(msb_tensor x, msb_tensor dy, float* dw, int cout, int cin, int fold_side, int n,
                        msb_dim3 dims, void* workspace, size_t workspace_bytes, void* stream);
/* out[c] (f32) += sum over n, voxels of x[n][c][v] for c < c_real (bias gradients; bf16 B8 view) */
int msb_channel_sum(msb_tensor x, int c_real, int n, int64_t s, float* out, void* stream);
/* debug switches for bring-up (key 0/1: swap LBO/SBO in the fwd / wgrad UMMA descriptors) */
int msb_debug_set(int key, int value);
/* key 5 = 1: the next msb_conv_k5(51)_fwd launches record, per CTA, the clocks their MMA-issuing warp spent
 * {in total, waiting for weight stages, waiting for halo tiles, waiting for a free accumulator}; synchronises. */
int msb_debug_read_prof(long long* host_out /* [148][4] */);

/* ---- fused Dice + cross-entropy loss ----------------------------------------------------------------
 * models/losses/dice_loss.py:76-102, cross_entropy_loss.py:47-87, loss_utils.py:31-40, mixes_losses.py:52-60. */
/* psum[C] (double) += sum over voxels of softmax(logits)[c]  (class_weights numerator/denominator) */
int msb_class_weight_sums(const float* logits, int n, int c, int64_t s, double* psum, void* stream);
int msb_class_weight_finalize(const double* psum, double count, int c, float* weights, void* stream);
/* acc (double) [3C+2] += { I_c, sum p_c^2, sum t_c, ce_num, ce_den } */
int msb_dice_ce_fwd(const float* logits, const int32_t* labels, const float* class_w, int n, int c, int64_t s,
                    int ignore_index, double* acc, void* stream);
/* result f32 [2+C] = { ce, dice_loss, per_channel_dice[C] } */
int msb_dice_ce_finalize(const double* acc, int c, float* result, void* stream);
/* The same three calls with DiceLoss's constructor options (medicalseg/models/losses/dice_loss.py:36-43,64-65):
 * dice_softmax != 0 normalises with softmax over the classes instead of the sigmoid (`sigmoid_norm=False`), dice_w
 * (device f32 [C] or NULL) multiplies the per-class intersections (`weight`).  The plain names are these with (0, NULL). */
int msb_dice_ce_fwd_ex(const float* logits, const int32_t* labels, const float* class_w, int n, int c, int64_t s,
                       int ignore_index, int dice_softmax, double* acc, void* stream);
int msb_dice_ce_finalize_ex(const double* acc, int c, const float* dice_w, float* result, void* stream);
int msb_dice_ce_bwd_ex(const float* logits, const int32_t* labels, const float* class_w, const double* acc, int n,
                       int c, int64_t s, int ignore_index, float coef_ce, float coef_dice, const float* coef_dev,
                       const float* dice_w, int dice_softmax, float* dlogits, void* stream);
/* Fused evaluation head (core/infer.py:79-92 argmax + core/val.py:101-118 loss behind out_tr.conv2, vnet.py:173-174):
 * logits = conv1x1(a) stay in registers.  a: B8 view (8, 16 or 32 channels, the first C live), w [C][C], b [C] or NULL.
 * Optional outputs (NULL = skip): pred int32 [N][S] (first maximum wins), acc double [3C+2] (+=, as msb_dice_ce_fwd;
 * needs labels + class_w), psum double [C] (+= softmax sums, as msb_class_weight_sums). */
int msb_eval_head(msb_tensor a, const float* w, const float* b, const int32_t* labels, const float* class_w, int n,
                  int c, int64_t s, int ignore_index, int32_t* pred, double* acc, double* psum, void* stream);
/* dlogits = coef_ce * dCE/dz + coef_dice * dDice/dz; coef_dev (device f32[2], may be NULL) multiplies the
 * two host coefficients so upstream gradients need no host synchronisation. */
int msb_dice_ce_bwd(const float* logits, const int32_t* labels, const float* class_w, const double* acc,
                    int n, int c, int64_t s, int ignore_index, float coef_ce, float coef_dice,
                    const float* coef_dev, float* dlogits, void* stream);

/* ---- optimizer.Momentum + L2 (cvlibs/config.py:212-214) on a flat parameter buffer ------------------
 * g' = grad_scale*g + wd*p;  v = mu*v + g';  p -= lr*v                                                 */
int msb_momentum_step(float* p, const float* g, float* v, int64_t count, float lr, float mu, float wd,
                      float grad_scale, void* stream);
/* same update with the learning rate read from DEVICE memory (one f32): lets a captured CUDA graph of the whole
 * train step be replayed while the PolynomialDecay schedule advances */
int msb_momentum_step_lrdev(float* p, const float* g, float* v, int64_t count, const float* lr_dev, float mu, float wd,
                            float grad_scale, void* stream);

/* ---- preprocessing (tools/preprocess_utils/values.py:54-87, geometry.py:31-69) -----------------------*/
int msb_hunorm(const float* src, float* dst, int64_t count, float hu_min, float hu_max, float hu_nan, void* stream);
/* minmax[2] f32 = {min, max} over src (NaNs ignored) */
int msb_minmax(const float* src, int64_t count, float* minmax, void* stream);
/* dst = clip((src - lo)/(hi - lo), 0, 1); if minmax != NULL lo/hi are read from device memory */
int msb_normalize(const float* src, float* dst, int64_t count, float lo, float hi, const float* minmax, void* stream);
/* scipy.ndimage.zoom(mode='nearest', order 0|1) with align-corners mapping.  pre_op: 0 none, 1 HUnorm
 * applied to every fetched sample before interpolation (fused HUnorm+resample), 2 normalize(lo,hi). */
int msb_resample_f32(const float* src, msb_dim3 in_dims, float* dst, msb_dim3 out_dims, int order, int pre_op,
                     float p0, float p1, float p2, void* stream);
int msb_resample_i32(const int32_t* src, msb_dim3 in_dims, int32_t* dst, msb_dim3 out_dims, void* stream);
int msb_label_remap(int32_t* labels, int64_t count, const int32_t* keys, const int32_t* vals, int nmap, void* stream);

/* paddle.argmax(logit, axis=1, keepdim=True, dtype='int32') of NCDHW f32 logits (medicalseg/core/infer.py:92): pred int32
 * [N][S], the first maximum wins. */
int msb_argmax_channels(const float* logits, int n, int c, int64_t s, int32_t* pred, void* stream);
/* nn.Dropout3D(p) masks of every dropout site of one forward (medicalseg/models/vnet.py:103,108,144-145,149-150):
 * out[i] = 0 or 1/(1-p), i over the concatenated [N][C] masks.  Counter-based generator keyed by (seed, *step_counter, i);
 * the kernel advances *step_counter (device memory), so a CUDA-graph replay draws new masks each step. */
int msb_dropout_masks(uint64_t seed, uint64_t* step_counter, float* out, int total, float p, void* stream);

/* ---- training augmentations on the device (medicalseg/transforms/functional.py:77-110, transform.py:46-72) ----
 * scipy.ndimage.rotate(axes=(axis_a, axis_b), reshape=False, mode='constant', order 0|1) of a [D][H][W] volume:
 * dst[o] = interp(src, M @ o + off) in every plane, cval outside 0 <= c <= n-1.  M = [[c, s], [-s, c]] with
 * scipy.special.cosdg/sindg of the angle and off = (n-1)/2 - M @ (n-1)/2, all f64 computed by the caller.  The i32
 * variant interpolates in f64 and rounds like SciPy (the reference rotates labels with order 1, transform.py:163-165). */
int msb_rotate3d_f32(const float* src, float* dst, msb_dim3 dims, int axis_a, int axis_b, double m00, double m01,
                     double m10, double m11, double off0, double off1, int order, float cval, void* stream);
int msb_rotate3d_i32(const int32_t* src, int32_t* dst, msb_dim3 dims, int axis_a, int axis_b, double m00, double m01,
                     double m10, double m11, double off0, double off1, int order, int32_t cval, void* stream);
/* np.flip(volume, axis) for 4-byte elements, out of place (functional.py:77-85) */
int msb_flip3d(const void* src, void* dst, msb_dim3 dims, int axis, void* stream);
/* Compose (transform.py:67-69): dst = src / max if max > 0 else src; minmax = device {min, max} from msb_minmax */
int msb_scale_by_max(const float* src, float* dst, int64_t count, const float* minmax, void* stream);

/* ---- deep-supervision heads (medicalseg/models/vnet_deepsup.py:257-272) ---------------------------------------
 * F.interpolate(x, size, mode='trilinear') with Paddle's defaults (align_corners=False, align_mode=0) on NCDHW f32:
 * per axis src = (in/out)*(dst+0.5)-0.5 clamped at 0.  nc = N*C planes.  _bwd is the exact adjoint (dsrc is
 * overwritten; deterministic gather, no atomics). */
int msb_trilinear_fwd(const float* src, int64_t nc, msb_dim3 in_dims, float* dst, msb_dim3 out_dims, void* stream);
int msb_trilinear_bwd(const float* ddst, int64_t nc, msb_dim3 out_dims, float* dsrc, msb_dim3 in_dims, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MEDSEG_B200_H_ */
