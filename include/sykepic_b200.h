/*
 * sykepic_b200 -- C ABI of the B200-native `sykepic prob` / `sykepic class` hot path.
 *
 * The reference (sykefi/syke-pic) is pure Python and has NO FFI / plugin interface
 * (SURVEY.md section 8b); its boundary for this path is a set of Python call
 * signatures.  Each entry point below names the reference interface it replaces
 * (file:line relative to the reference checkout).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch/Python types.
 *   - Every function returns an int status: 0 = SPK_OK, negative = error.
 *     `spk_last_error(ctx)` gives a human-readable message for the last failure on
 *     that context (or the last context-free failure of the calling thread when
 *     ctx == NULL).
 *   - The CALLER owns every buffer it passes (device pointers from e.g.
 *     torch.Tensor.data_ptr(), pinned host buffers).  The library owns only what
 *     hangs off the opaque `spk_ctx` (packed/folded weights, activation workspace,
 *     TMA descriptors) created by spk_create and freed by spk_destroy.
 *   - All device work is enqueued asynchronously on the context's CUDA stream
 *     (spk_create / spk_set_stream); nothing synchronises unless stated.
 *   - One context is used by one host thread at a time (one process per GPU).
 */
#ifndef SYKEPIC_B200_H
#define SYKEPIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPK_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------ */
#define SPK_OK                 0
#define SPK_ERR_INVALID       -1  /* bad argument */
#define SPK_ERR_CUDA          -2  /* CUDA runtime / driver error (message has the detail) */
#define SPK_ERR_FAULTY_BIN    -3  /* ROI slice runs past the .roi bytes: the reference's ValueError
                                     "Faulty raw data" (sykepic/compute/probability.py:111-112); skip the bin */
#define SPK_ERR_EMPTY_RESIZE  -4  /* aspect ratio > T:1 gives a 0-pixel side: cv2.error in the reference,
                                     bin skipped (probability.py:113-114) */
#define SPK_ERR_PARSE         -5  /* malformed .adc line (reference: IndexError / ValueError -> bin skipped) */
#define SPK_ERR_CAPACITY      -6  /* caller-provided output buffer too small */
#define SPK_ERR_UNSUPPORTED   -7
#define SPK_ERR_STATE         -8  /* call order violated (e.g. spk_forward before spk_net_end) */
#define SPK_ERR_IO            -9  /* a file could not be opened / read / written (message has errno's text); the
                                     reference raises OSError and logs "Unexpected error" for the bin */

/* ---- enums -------------------------------------------------------------------------- */
enum { SPK_BORDER_MODE = 0, SPK_BORDER_BLACK = 1, SPK_BORDER_WHITE = 2 }; /* sykepic/train/image.py:20-28 */
enum { SPK_DTYPE_F32 = 0, SPK_DTYPE_BF16 = 1, SPK_DTYPE_U8 = 2,
       SPK_DTYPE_SPLIT = 3 /* internal activation format of the FP32_TC precision: fp16 hi | fp16 lo in one 32-bit word */ };
enum { SPK_LAYOUT_NCHW = 0, SPK_LAYOUT_NHWC = 1 };
enum { SPK_PRECISION_FP32 = 0, SPK_PRECISION_BF16 = 1,
       SPK_PRECISION_FP32_TC = 2 /* fp32-level accuracy on the bf16 tensor cores (activations and weights as bf16 hi + lo;
                                    ~1e-5 relative per layer from the tensor core's truncating fp32 accumulation) */ };
/* implementation selector for convolutions in bf16 precision: AUTO / TCGEN05 pick the halo-resident
 * kernel for 3x3 stride-1 layers on large maps and the tap-per-TMA kernel elsewhere; TCGEN05_TAPS
 * forces the tap-per-TMA kernel everywhere (A/B comparison). */
enum { SPK_CONV_AUTO = 0, SPK_CONV_SIMT = 1, SPK_CONV_TCGEN05 = 2, SPK_CONV_TCGEN05_TAPS = 3 };

typedef struct spk_ctx spk_ctx;

/* ---- context ------------------------------------------------------------------------ */
int spk_abi_version(void);
/* stream: a cudaStream_t (NULL = the legacy default stream). */
int spk_create(int device, void* stream, spk_ctx** out);
int spk_destroy(spk_ctx* ctx);
int spk_set_stream(spk_ctx* ctx, void* stream);
int spk_synchronize(spk_ctx* ctx);
const char* spk_last_error(const spk_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t spk_launch_count(const spk_ctx* ctx);

/* ---- per-kernel timing (bench.py's roofline leg) -------------------------------------------------
 * While profiling is on, every launch of spk_preprocess / spk_forward is bracketed by CUDA events on
 * the context's stream.  spk_profile_read synchronises, sums the elapsed time, the algorithmic FLOPs
 * and the algorithmic bytes of the launches recorded since spk_profile_begin per category
 * (SPK_PROF_*), and, when `detail` is not NULL, writes one text line per launch
 * ("<category> <ms> <flops> <bytes> <description>\n") into it.  Profiling is off by default and
 * costs nothing then. */
enum { SPK_PROF_PREPROCESS = 0, SPK_PROF_CONV_TC = 1, SPK_PROF_CONV_SIMT = 2, SPK_PROF_STEM = 3,
       SPK_PROF_POOL = 4, SPK_PROF_BN_RELU = 5, SPK_PROF_HEAD = 6, SPK_PROF_CATEGORIES = 8 };
/* SPK_PROFILE_EVENTS (default): CUDA events around every launch -- simple, but an event between two launches serialises
 * them (no programmatic-dependent-launch overlap) and adds a few microseconds each, so the per-kernel times sum to MORE
 * than the step.  SPK_PROFILE_STAMPS: every tcgen05 kernel, the stem and the head stamp their earliest CTA start / latest
 * CTA end on the GPU's global timer into a context-owned buffer; the launches stay back to back as in production.
 * spk_profile_read then reports per launch the IN-STEP time = end - max(previous launch's end, own start), which sums to
 * the step, and in the detail line the raw span (" |span_ms=..."); kernels without stamps (pools, CUDA-core convolutions,
 * K1) are folded into the next stamped launch. */
enum { SPK_PROFILE_EVENTS = 0, SPK_PROFILE_STAMPS = 1 };
int spk_profile_mode(spk_ctx* ctx, int mode);
int spk_profile_begin(spk_ctx* ctx);
int spk_profile_read(spk_ctx* ctx, double ms[SPK_PROF_CATEGORIES], double flops[SPK_PROF_CATEGORIES],
                     double bytes[SPK_PROF_CATEGORIES], int64_t launches[SPK_PROF_CATEGORIES],
                     char* detail, int64_t detail_cap);
int spk_profile_end(spk_ctx* ctx);

/* ---- A2: .adc parsing (host) ---------------------------------------------------------
 * Replaces the per-line parsing in sykepic/utils/ifcb.py:101-110 (raw_to_png) and
 * :133-145 (next_roi): ROI id = 1-based LINE number, width/height/start = comma fields
 * 15/16/17 with Python int() syntax (optional whitespace and sign), rows with
 * width < 1 or height < 1 skipped.  Universal newlines (\n, \r\n, \r).
 * Outputs hold the non-empty ROIs in file order; *n_out their count, *n_lines the
 * number of lines seen.  SPK_ERR_PARSE on a short / non-integer line (the reference
 * raises and the bin is skipped), SPK_ERR_CAPACITY when cap is too small. */
int spk_adc_parse(const char* text, int64_t len, int64_t cap,
                  int32_t* roi_id, int32_t* width, int32_t* height, int64_t* start,
                  int64_t* n_out, int64_t* n_lines);

/* Geometry checks the reference performs implicitly while decoding / resizing one bin:
 * start + w*h <= roi_len for every ROI (else SPK_ERR_FAULTY_BIN, ifcb.py:111-116 reshape)
 * and both resized sides >= 1 px for the target (else SPK_ERR_EMPTY_RESIZE,
 * sykepic/train/image.py:183-198 + cv2.resize).  *first_bad = index of the first
 * offending ROI (or -1). */
int spk_rois_validate(const int32_t* width, const int32_t* height, const int64_t* start, int64_t n,
                      int64_t roi_len, int target_h, int target_w, int64_t* first_bad);

/* One bin from disk in a single call (the loader threads of the host pipeline spend their time here, outside the
 * Python interpreter lock): read and parse `adc_path` (as spk_adc_parse), read `roi_path` straight into `roi_buf`
 * (typically pinned host memory; *roi_len = the file's size, SPK_ERR_CAPACITY if it exceeds roi_cap -- *roi_len is
 * set, retry with a larger buffer), then check the geometry (as spk_rois_validate).  Replaces the file reads of
 * ifcb.raw_to_png (sykepic/utils/ifcb.py:76-118: `open(adc)`, `np.fromfile(roi)`).  SPK_ERR_IO when a file cannot
 * be read. */
int spk_bin_load(const char* adc_path, const char* roi_path, uint8_t* roi_buf, int64_t roi_cap, int64_t desc_cap,
                 int32_t* roi_id, int32_t* width, int32_t* height, int64_t* start, int64_t* n_out, int64_t* roi_len,
                 int target_h, int target_w);

/* get_new_dims of sykepic/train/image.py:183-198 (float64 arithmetic, truncation). */
void spk_new_dims(int h, int w, int target_h, int target_w, int* new_h, int* new_w);

/* ---- A2+A3+A4+A5(loader): ROI decode + eval transform (device, kernel K1) -------------
 * Replaces ifcb.raw_to_png (utils/ifcb.py:76-118) + ImageDataset.__getitem__
 * (train/data.py:210-231) + Compose.__call__ (train/image.py:25-56: mode border,
 * get_new_dims, cv2 INTER_LINEAR fixed-point resize incl. the 2x INTER_AREA switch and
 * the identity copy, centred copyMakeBorder) + ToTensor (+ optional Normalize through
 * the LUT) + default collate, for a batch of n ROIs.
 *   roi_bytes   device, the raw .roi byte stream of the bin (roi_len bytes)
 *   start/w/h   device, per-ROI start byte (int64), width, height (int32)
 *   channels    1 or 3 (3 = identical planes unless the LUT differs per channel)
 *   out_dtype   SPK_DTYPE_F32 / SPK_DTYPE_BF16 (LUT applied) or SPK_DTYPE_U8 (the
 *               resized + padded bytes, no LUT, channels must be 1)
 *   lut         device float[3*256] or NULL (NULL = ToTensor's v/255 for every channel)
 *   out         device, [n, channels, T_h, T_w] (NCHW) or [n, T_h, T_w, channels] (NHWC)
 * ROIs whose geometry is invalid (see spk_rois_validate) are filled with the border
 * value / zero and counted in the context's device-side fault counter, which
 * spk_fault_count reads back (synchronises). */
int spk_preprocess(spk_ctx* ctx, const uint8_t* roi_bytes, int64_t roi_len,
                   const int64_t* start, const int32_t* width, const int32_t* height, int64_t n,
                   int target_h, int target_w, int border_mode, int channels,
                   int out_dtype, int out_layout, const float* lut, void* out);
/* The fault counter also counts, in SPK_PRECISION_FP32_TC, every tile of a convolution or of the stem whose output left the
 * fp16 range of the split activation format (|x| > 65504: the value saturates); a caller must treat a forward pass after
 * which the counter moved as invalid (sykepic_b200/engine.py raises ArithmeticError and points at --precision fp32). */
int spk_fault_count(spk_ctx* ctx, int64_t* count);

/* ---- A5/A6/A7: the network (kernels K2) -----------------------------------------------
 * Replaces prepare_model's get_network + load_state_dict (compute/probability.py:118-130,
 * train/config.py:63-77, train/network.py:14-64) and TorchVisionNet.forward's `base`
 * (train/network.py:66-68).  The host walks the state_dict (the checkpoint format is
 * kept) and describes the graph op by op; the library folds eval-mode BatchNorm into
 * the convolution (w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps), in double),
 * packs weights for its kernels and plans the activation workspace.
 * Activations are NHWC.  Buffer ids are small integers chosen by the caller;
 * buffer 0 is the network input (the preprocessed batch, set at spk_forward). */
int spk_net_begin(spk_ctx* ctx, int target_h, int target_w, int in_channels, int precision, int max_batch);
/* Declare activation buffer `buf` (NHWC, `channels` wide).  Optional for buffers a single op writes
 * completely (their shape is inferred); required for concatenation buffers that several convolutions
 * fill slice by slice (DenseNet blocks, torch.cat in torchvision's _DenseLayer). */
int spk_net_buffer(spk_ctx* ctx, int buf, int h, int w, int channels);
/* weight: host float[cout*cin*kh*kw] (OIHW, as in the state_dict); bn_*: host float[cout] or NULL;
 * bias: host float[cout] or NULL; res_buf: buffer added before the ReLU, or -1.
 * The convolution reads channels [in_c_off, in_c_off+cin) of in_buf and writes channels
 * [out_c_off, out_c_off+cout) of out_buf.
 * When cin is 3 and the network input has 1 channel (identical planes), the three
 * input-channel slices are summed (exact fold of R=G=B). */
int spk_net_conv(spk_ctx* ctx, int in_buf, int in_c_off, int out_buf, int out_c_off, int res_buf,
                 const float* weight, int cout, int cin, int kh, int kw, int stride, int pad,
                 const float* bn_gamma, const float* bn_beta, const float* bn_mean, const float* bn_var, float bn_eps,
                 const float* bias, int relu, int impl);
int spk_net_maxpool(spk_ctx* ctx, int in_buf, int out_buf, int k, int stride, int pad);
/* k x k / stride average pool, no padding (DenseNet transition). */
int spk_net_avgpool(spk_ctx* ctx, int in_buf, int out_buf, int k, int stride);
/* Eval-mode BatchNorm (+ ReLU) as a stand-alone op on channels [0, channels) of in_buf: DenseNet's
 * pre-activation norm1 / transition norm / norm5, which cannot be folded into the producer. */
int spk_net_bn_relu(spk_ctx* ctx, int in_buf, int out_buf, int channels,
                    const float* bn_gamma, const float* bn_beta, const float* bn_mean, const float* bn_var, float bn_eps,
                    int relu);
/* Global average pool of `in_buf` followed by the all-Linear head (train/network.py:56-63,
 * 68-69; no activations, Dropout = identity in eval).  weights[i]: host float[dims[i+1]*dims[i]],
 * biases[i]: host float[dims[i+1]].  The chain is folded into one affine map in double. */
int spk_net_head(spk_ctx* ctx, int in_buf, int n_layers, const float* const* weights,
                 const float* const* biases, const int* dims);
int spk_net_end(spk_ctx* ctx);
/* Bytes of device memory the network holds (weights + workspace). */
int64_t spk_net_bytes(const spk_ctx* ctx);
/* Convolution layers of a BF16 / FP32_TC network that SPK_CONV_AUTO had to place on the CUDA-core kernel because no
 * tcgen05 kernel takes their geometry (0 for every torchvision ResNet / DenseNet; each one is also reported on stderr). */
int spk_net_simt_layers(const spk_ctx* ctx);

/* ---- A6/A8/A10: forward + softmax + per-class threshold (kernels K2, K3) ---------------
 * Replaces net_pass (compute/probability.py:180-197: logits * ln(1.3) in fp32, softmax)
 * and, fused, prediction.row_prediction (compute/prediction.py:49-71) evaluated on the
 * 5-decimal values the CSV would hold.
 *   x            device, preprocessed batch in the layout spk_net_begin implies:
 *                [n, T, T] u8 (bf16 precision, 1 channel) or [n, T, T, C] f32 NHWC
 *   softmax_scale  ln(1.3) in the reference (probability.py:18,192-193); 0 = plain softmax
 *   thr_q        device int32[K] or NULL: per class, the smallest value of
 *                round(p * 1e5) that counts as "above threshold" (INT32_MAX = class has
 *                no threshold); see spk_threshold_quantize.
 *   probs        device float[n*K]
 *   label        device int32[n] or NULL, classified device uint8[n] or NULL
 */
int spk_forward(spk_ctx* ctx, const void* x, int64_t n, float softmax_scale, const int32_t* thr_q,
                float* probs, int32_t* label, uint8_t* classified);
/* Logits of the last spk_forward (device float[n*K], owned by the context). */
const float* spk_last_logits(const spk_ctx* ctx);
/* Debug / test tap: copy activation buffer `buf` of the last forward to host as fp32 NHWC. */
int spk_net_read_buffer(spk_ctx* ctx, int buf, int64_t n, float* host_out, int64_t cap_elems, int* h, int* w, int* c);

/* Host: turn a decimal threshold into the integer domain of round(p*1e5):
 * strict = 0: smallest q with (double)(q/1e5 as parsed from "%.5f") >= thr   (dict thresholds, prediction.py:63-64)
 * strict = 1: smallest q with value > thr                                    (scalar threshold, prediction.py:58-59) */
int32_t spk_threshold_quantize(double thr, int strict);

/* ---- A10: reading a `.prob.csv` back (host) ------------------------------------------------
 * `sykepic class` and every downstream tool re-read the probability files with pandas.read_csv inside
 * prediction_dataframe (compute/prediction.py:8-28); at 5000 x 50 values per bin that read is 3/4 of the time of
 * `sykepic class`.  These two calls parse the text as written by spk_format_prob_csv / probabilities_to_csv:
 * a header line, then "<integer>,<decimal>,...".  Decimals are converted exactly as pandas' default parser
 * does for them (correctly rounded: digits / 10^k in one IEEE division; plain decimals of up to 15 digits only).
 *   spk_prob_csv_shape   rows (blank lines skipped) and value columns (header fields - 1)
 *   spk_prob_csv_parse   roi[n_rows], values[n_rows * n_cols] (row-major doubles)
 * SPK_ERR_PARSE on anything that is not this layout (quotes, ragged rows, non-numeric tokens): the caller then
 * falls back to pandas, which decides as the reference would. */
int spk_prob_csv_shape(const char* text, int64_t len, int64_t* n_rows, int* n_cols);
int spk_prob_csv_parse(const char* text, int64_t len, int64_t n_rows, int n_cols, int64_t* roi, double* values);

/* ---- image mode: PNG scanline filters (host) ----------------------------------------------
 * `sykepic prob --image-dir / --images` (compute/probability.py:28-36,165-177) reads ROI
 * images with cv2.imread (train/data.py:217-219).  Here the PNG container and zlib stream are
 * handled by the Python host; this reverses the five scanline filters (None, Sub, Up, Average,
 * Paeth) of the inflated IDAT data, the only per-byte sequential part of the decode.
 *   raw     h rows of (1 filter-type byte + stride bytes)
 *   bpp     bytes per pixel (1 gray, 2 gray+alpha, 3 RGB, 4 RGBA; 8-bit samples)
 *   out     h * stride bytes, the unfiltered samples
 * SPK_ERR_PARSE on an unknown filter type. */
int spk_png_unfilter(const uint8_t* raw, int64_t h, int64_t stride, int bpp, uint8_t* out);

/* A whole sample's ROI images at once, on `threads` host threads (0 = all cores, at most 16), straight into one byte
 * stream laid out like a `.roi` file -- what the DataLoader workers' cv2.imread calls produce one by one in the reference
 * (train/data.py:217-219).  8-bit gray / gray+alpha / RGB(A) with equal colour channels, non-interlaced; chunk CRCs are
 * checked (libpng rejects such files too).
 *   spk_png_probe          width / height of every file (reads the 33-byte IHDR prefix only)
 *   spk_png_decode_batch   file i -> out[start[i] .. start[i] + width[i] * height[i]) as one gray plane
 * SPK_ERR_PARSE with the path and reason in spk_last_error(NULL), *first_bad = index of the first offending file. */
int spk_png_probe(const char* const* paths, int64_t n, int32_t* width, int32_t* height, int threads, int64_t* first_bad);
int spk_png_decode_batch(const char* const* paths, int64_t n, const int32_t* width, const int32_t* height, const int64_t* start,
                         uint8_t* out, int64_t out_len, int threads, int64_t* first_bad);

/* ---- A9: .prob.csv formatting (host) ----------------------------------------------------
 * Replaces probabilities_to_csv (compute/probability.py:200-206): header line, then per
 * ROI "<id>,<p:.5f>,...\n" with p widened to double and correctly rounded.
 * Writes at most cap bytes into out, *len = bytes needed. */
int spk_format_prob_csv(const char* header_line, const int32_t* roi_id, const float* probs,
                        int64_t n, int k, char* out, int64_t cap, int64_t* len);
/* The same text written to `path` (created / truncated; parent directories must exist) in one call: what
 * probabilities_to_csv's `open(csv_path, "w")` + `write` do (probability.py:205-206).  SPK_ERR_IO on failure. */
int spk_prob_csv_write(const char* path, const char* header_line, const int32_t* roi_id, const float* probs,
                       int64_t n, int k, int64_t* bytes_written);

#ifdef __cplusplus
}
#endif
#endif /* SYKEPIC_B200_H */
