/* nncabac -- C ABI of the host-side NNC entropy coder and dependent quantiser (libnncabac.so).
 *
 * Replaces the pybind11 module `deepCABAC` the reference binds (upstream fraunhoferhhi/nncodec,
 * extensions/deepCABAC; absent from /root/reference, no version pinned, not installable offline).  Call sites in the
 * reference checkout jihyounchoi/vanilla-nerf-model-compression-using-lsa-enhanced-nncodec:
 *   nnc_core/approximator/baseline.py:24-57   Encoder(), initCtxModels, quantLayer
 *   nnc_core/approximator/baseline.py:89-98   Decoder(), dequantLayer
 *   nnc_core/coder/baseline.py:5-57           iae_v, initCtxModels, encodeLayer / decodeLayer / decodeLayerAndCreateEPs
 *   nnc_core/coder/__init__.py:118-140        Encoder.finish(), Decoder.setStream
 *   nnc_core/coder/__init__.py:439-483        Decoder.setEntryPoints, Decoder.finish() -> bytes read
 * The Python class layer with deepCABAC's method names lives in nerfq_b200/deepcabac.py (ctypes over this ABI).
 *
 * This is sequential host code by design (north_star: "DeepCABAC entropy coding stays on the reference's sequential
 * coder"); the data-parallel part of quantisation -- uniform reconstruction quantisation, dq_flag = 0 -- is the CUDA kernel
 * nerfq_quantize_urq (include/nerfq.h).  PARITY UNPINNED against the real deepCABAC: the algorithms restate ISO/IEC
 * 15938-17 (NNC) -- context-adaptive binary arithmetic coding of sig / sign / greater-than-x / Exp-Golomb remainder bins,
 * 8-state dependent (trellis-coded) quantisation -- and are checked for self-consistency (encode -> decode identity,
 * exact byte accounting, the unmodified reference's compress -> decompress round trip).
 *
 * Conventions: all pointers are HOST pointers; return 0 ok, -1 bad argument, -2 unsupported (scan_order > 0),
 * -3 corrupt / exhausted stream; handles are single-threaded.
 */
#ifndef NNCABAC_H
#define NNCABAC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nncabac_encoder nncabac_encoder;
typedef struct nncabac_decoder nncabac_decoder;

/* ---- encoder (deepCABAC.Encoder) ---- */
nncabac_encoder* nncabac_encoder_new(void);
void nncabac_encoder_free(nncabac_encoder* e);
/* initCtxModels(cabac_unary_length_minus1, param_opt_flag): resets every context model and the neighbour state. */
int nncabac_encoder_init_ctx(nncabac_encoder* e, int cabac_unary_length_minus1, int param_opt_flag);
/* iae_v(n_bits, value): n_bits-wide two's-complement integer as bypass bins. */
int nncabac_encoder_iae_v(nncabac_encoder* e, int n_bits, int value);
/* quantLayer: w[n] -> lvl[n]; returns the qp actually used through *qp_used (raised until the largest level fits int32,
 * baseline.py:60-62).  dq_flag 0: nearest-integer rounding of |w|/delta (ties away from zero) -- same arithmetic as the
 * CUDA kernel; dq_flag 1: 8-state trellis search (Viterbi, squared error + lambda_scale * delta^2 * estimated bits),
 * output = integers on the union grid, value = lvl * delta for BOTH modes (what dequantLayer applies). */
int nncabac_quant_layer(nncabac_encoder* e, const float* w, int32_t* lvl, int64_t n, int dq_flag, int qp_density, int qp,
                        float lambda_scale, int cabac_unary_length_minus1, int scan_order, int* qp_used);
/* encodeLayer(levels, dq_flag, scan_order): row-major scan (scan_order 0; block scans are not implemented: -2). */
int nncabac_encoder_encode_layer(nncabac_encoder* e, const int32_t* lvl, int64_t n, int dq_flag, int scan_order);
/* finish(): terminates the arithmetic codeword; *data stays valid until the encoder is freed or reused. */
int nncabac_encoder_finish(nncabac_encoder* e, const uint8_t** data, int64_t* size);

/* ---- decoder (deepCABAC.Decoder) ---- */
nncabac_decoder* nncabac_decoder_new(void);
void nncabac_decoder_free(nncabac_decoder* d);
/* setStream(bytes): the bytes are copied. */
int nncabac_decoder_set_stream(nncabac_decoder* d, const uint8_t* data, int64_t size);
int nncabac_decoder_init_ctx(nncabac_decoder* d, int cabac_unary_length_minus1);
int nncabac_decoder_iae_v(nncabac_decoder* d, int n_bits, int* value);
int nncabac_decoder_decode_layer(nncabac_decoder* d, int32_t* lvl, int64_t n, int dq_flag, int scan_order);
/* finish(): reads the terminating bin and returns the number of bytes the codeword occupied. */
int nncabac_decoder_finish(nncabac_decoder* d, int64_t* bytes_read);
/* dequantLayer: out[i] = (float)lvl[i] * delta(qp, qp_density)  (nnc_core/approximator/codebook.py:346-356). */
int nncabac_dequant_layer(float* out, const int32_t* lvl, int64_t n, int qp_density, int qp);

/* delta(qp, qp_density), nnc_core/common.py:28-46 */
float nncabac_stepsize(int qp, int qp_density);

#ifdef __cplusplus
}
#endif
#endif /* NNCABAC_H */
