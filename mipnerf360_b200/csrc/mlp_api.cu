// MLP-level entry points: a whole proposal / NeRF MLP (model.py:43-53, :131-158) forward or backward as one C call
// that chains the tcgen05 GEMM launches of gemm_tcgen05.cu.  No arithmetic of its own.
#include "common.cuh"

using namespace mip360;

static int check_layers(const mip360_layer* trunk, int n_trunk, const mip360_layer* head, const char* who) {
  MIP_REQUIRE(trunk && head && n_trunk >= 1 && n_trunk <= 32, "%s: bad layer table", who);
  int k = trunk[0].k_pad;
  for (int l = 0; l < n_trunk; ++l) {
    MIP_REQUIRE(trunk[l].W && trunk[l].bias, "%s: layer %d has null weights", who, l);
    MIP_REQUIRE(trunk[l].k_pad == k, "%s: layer %d expects %d inputs, previous layer produces %d", who, l, trunk[l].k_pad, k);
    k = trunk[l].n_pad;
  }
  MIP_REQUIRE(head->W && head->bias && head->k_pad == k && head->n_pad == 64, "%s: head must be [64, %d] (zero padded)", who, k);
  return MIP360_OK;
}

extern "C" {

int mip360_mlp_fwd(const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk, const mip360_layer* head,
                   int n_valid, uint16_t* const* acts, int n_act_bufs, float* out, mip360_stream_t stream) {
  int rc = check_layers(trunk, n_trunk, head, "mlp_fwd");
  if (rc != MIP360_OK) return rc;
  MIP_REQUIRE(acts && out && (n_act_bufs == n_trunk || n_act_bufs == 2), "mlp_fwd: need %d (saved) or 2 (ping-pong) buffers", n_trunk);
  if (M <= 0) return MIP360_OK;
  MIP_REQUIRE(x, "mlp_fwd: null input");
  const uint16_t* h = x;
  for (int l = 0; l < n_trunk; ++l) {
    uint16_t* y = acts[n_act_bufs == 2 ? (l & 1) : l];
    MIP_REQUIRE(y, "mlp_fwd: activation buffer %d is null", l);
    rc = mip360_linear_fwd(h, trunk[l].W, trunk[l].bias, M, trunk[l].n_pad, trunk[l].k_pad, trunk[l].act, y, nullptr, 0, stream);
    if (rc != MIP360_OK) return rc;
    h = y;
  }
  return mip360_linear_fwd(h, head->W, head->bias, M, 64, head->k_pad, head->act, nullptr, out, n_valid, stream);
}

int mip360_mlp_fwd_fused_head(const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk, const float* head_w4,
                              uint16_t* const* acts, int n_act_bufs, float* out, mip360_stream_t stream) {
  MIP_REQUIRE(trunk && n_trunk >= 1 && n_trunk <= 32 && head_w4, "mlp_fwd_fused_head: bad layer table");
  MIP_REQUIRE(acts && out && (n_act_bufs == n_trunk || n_act_bufs == 2), "mlp_fwd_fused_head: need %d (saved) or 2 (ping-pong) buffers", n_trunk);
  if (M <= 0) return MIP360_OK;
  MIP_REQUIRE(x, "mlp_fwd_fused_head: null input");
  MIP_CUDA(cudaMemsetAsync(out, 0, (size_t)M * 4 * sizeof(float), (cudaStream_t)stream));
  const uint16_t* h = x;
  int rc;
  for (int l = 0; l < n_trunk; ++l) {
    MIP_REQUIRE(trunk[l].W && trunk[l].bias, "mlp_fwd_fused_head: layer %d has null weights", l);
    uint16_t* y = acts[n_act_bufs == 2 ? (l & 1) : l];
    if (l + 1 < n_trunk) {
      MIP_REQUIRE(y, "mlp_fwd_fused_head: activation buffer %d is null", l);
      rc = mip360_linear_fwd(h, trunk[l].W, trunk[l].bias, M, trunk[l].n_pad, trunk[l].k_pad, trunk[l].act, y, nullptr, 0, stream);
    } else {
      // inference (ping-pong buffers): nobody reads the last trunk activation, so it is not written
      rc = mip360_linear_fwd_head(h, trunk[l].W, trunk[l].bias, M, trunk[l].n_pad, trunk[l].k_pad, trunk[l].act,
                                  n_act_bufs == 2 ? nullptr : y, head_w4, out, stream);
    }
    if (rc != MIP360_OK) return rc;
    h = y;
  }
  return MIP360_OK;
}

int mip360_mlp_bwd_fused_head(const float* g_out, const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk,
                              const float* head_w4, uint16_t* const* acts, float* const* dW, float* const* db,
                              uint16_t* dz0, uint16_t* dz1, mip360_stream_t stream) {
  MIP_REQUIRE(trunk && n_trunk >= 1 && n_trunk <= 32 && head_w4 && acts && dW && db && dz0 && dz1, "mlp_bwd_fused_head: null pointer");
  if (M <= 0) return MIP360_OK;
  MIP_REQUIRE(g_out && x, "mlp_bwd_fused_head: null input");
  const int L = n_trunk;
  uint16_t* dz = dz0;
  // head: gradient entering the last trunk layer, head weight and bias gradients, one pass over acts[L-1]
  int rc = mip360_head_bwd(g_out, head_w4, acts[L - 1], M, trunk[L - 1].n_pad, trunk[L - 1].act, dz, dW[L], trunk[L - 1].n_pad,
                           db[L], stream);
  if (rc != MIP360_OK) return rc;
  for (int l = L - 1; l >= 0; --l) {  // trunk layer l maps (l == 0 ? x : acts[l-1]) -> acts[l]
    const uint16_t* in = l == 0 ? x : acts[l - 1];
    rc = mip360_linear_wgrad(dz, in, M, trunk[l].n_pad, trunk[l].k_pad, dW[l], db[l], stream);
    if (rc != MIP360_OK) return rc;
    if (l > 0) {
      MIP_REQUIRE(trunk[l].Wt, "mlp_bwd_fused_head: layer %d has no transposed weights", l);
      uint16_t* nxt = (dz == dz0) ? dz1 : dz0;
      rc = mip360_linear_dgrad(dz, trunk[l].Wt, acts[l - 1], M, trunk[l].n_pad, trunk[l].k_pad, trunk[l - 1].act, nxt, stream);
      if (rc != MIP360_OK) return rc;
      dz = nxt;
    }
  }
  return MIP360_OK;
}

int mip360_mlp_bwd(const float* g_out, const float* out, const uint16_t* x, int M, const mip360_layer* trunk, int n_trunk,
                   const mip360_layer* head, int n_valid, uint16_t* const* acts, float* const* dW, float* const* db,
                   uint16_t* dz_head, uint16_t* dz0, uint16_t* dz1, mip360_stream_t stream) {
  int rc = check_layers(trunk, n_trunk, head, "mlp_bwd");
  if (rc != MIP360_OK) return rc;
  MIP_REQUIRE(acts && dW && db && dz_head && dz0 && dz1 && head->Wt, "mlp_bwd: null pointer");
  if (M <= 0) return MIP360_OK;
  MIP_REQUIRE(g_out && x && (head->act != 2 || out), "mlp_bwd: null input");
  const int L = n_trunk;
  // head: gradient of the head pre-activations, bf16, padded to 64 columns
  rc = mip360_head_grad_pack(g_out, out, M, n_valid, head->act, dz_head, stream);
  if (rc != MIP360_OK) return rc;
  rc = mip360_linear_wgrad(dz_head, acts[L - 1], M, 64, head->k_pad, dW[L], db[L], stream);
  if (rc != MIP360_OK) return rc;
  // into the trunk: derivative of the last trunk activation from its saved output
  uint16_t* dz = dz0;
  rc = mip360_linear_dgrad(dz_head, head->Wt, acts[L - 1], M, 64, head->k_pad, trunk[L - 1].act, dz, stream);
  if (rc != MIP360_OK) return rc;
  for (int l = L - 1; l >= 0; --l) {  // trunk layer l maps (l == 0 ? x : acts[l-1]) -> acts[l]
    const uint16_t* in = l == 0 ? x : acts[l - 1];
    rc = mip360_linear_wgrad(dz, in, M, trunk[l].n_pad, trunk[l].k_pad, dW[l], db[l], stream);
    if (rc != MIP360_OK) return rc;
    if (l > 0) {
      MIP_REQUIRE(trunk[l].Wt, "mlp_bwd: layer %d has no transposed weights", l);
      uint16_t* nxt = (dz == dz0) ? dz1 : dz0;
      rc = mip360_linear_dgrad(dz, trunk[l].Wt, acts[l - 1], M, trunk[l].n_pad, trunk[l].k_pad, trunk[l - 1].act, nxt, stream);
      if (rc != MIP360_OK) return rc;
      dz = nxt;
    }
  }
  return MIP360_OK;
}

}  // extern "C"
