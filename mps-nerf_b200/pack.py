"""Pack the live weights into the blob consumed by the tcgen05 kernels (csrc/dense_tc.cu).

Layout (bytes; every weight tile is K-major bf16 in SWIZZLE_128B chunks, see
``engine.pack_kmajor_sw128``), in the exact order the kernels stream it:

  transformer, per layer l (466 944 B):
      qkv_0, qkv_1, qkv_2 [3 chunks x 192 rows each]; Wo_0 [1 x 160]; qkv_3; Wo_1; Wo_2; Wo_3;
      W1 rows 0..63 [3 x 64], W1 rows 64..127 [3 x 64]; W2[2 x 160]
      qkv_h rows = (q_h | k_h | v_h) = to_qkv rows 64h.., 256+64h.., 512+64h..   (K = 155 -> 192)
      LayerNorm affine folded in: columns scaled by gamma, K column 155 = W beta (against a constant 1.0 the
      LayerNorm epilogue writes there); W1 scaled by gamma, W1 beta added to b1
  MLP (1 441 792 B = 44 ring slots of 2 chunks of 128 rows x 128 B), every 256-output layer split
      into its two N-halves (output rows 0..127, then 128..255), each half 4 chunks (K = 256) per K part:
      L0[h0 | h1] (K order: tok0 155, 5 zero, PE 39, zero to 256); L1..L4[h0 | h1] each;
      L5[h0: x part, h part | h1: x part, h part]; L6, L7, feature[h0 | h1];
      views[tok1 part 4 chunks | feature part 4 chunks] (128 outputs)
  fp32 section:
      per transformer layer 1088 floats: ln1_g, ln1_b, pend_in, ln2_g, ln2_b, pend_mid (160 each), b1 (128)
      then pend_out (160): the biases of to_out / net.3 are *deferred*: the residual stream in
      TMEM never contains them, every reader adds the running sum (pend_*) instead;
      MLP: b0..b7 (8 x 256), w_alpha (256), b_feature (256), b_views (128), w_rgb (3 x 128),
      b_alpha, b_rgb (3).
"""
import torch

from .engine import pack_kmajor_sw128

T_LAYER_BYTES = 12 * 24576 + 4 * 20480 + 3 * 16384 + 2 * 20480
T_BYTES = 2 * T_LAYER_BYTES
M_BYTES = 44 * 32768
T_FLOATS = 2 * 1088 + 160
M_FLOATS = 8 * 256 + 256 + 256 + 128 + 384 + 4
FLOAT_OFFSET = T_BYTES + M_BYTES
BLOB_BYTES = FLOAT_OFFSET + 4 * (T_FLOATS + M_FLOATS)


def _pad160(v):
    out = torch.zeros(160, dtype=torch.float32, device=v.device)
    out[:v.numel()] = v.float()
    return out


@torch.no_grad()
def pack_weights_bf16(net, device=None):
    sd = {k: v.detach().float() for k, v in net.state_dict().items()}
    dev = device or sd["alpha_linear.weight"].device
    sd = {k: v.to(dev) for k, v in sd.items() if k.startswith(("transformer.", "pts_linears.", "alpha_linear.",
                                                               "feature_linear.", "views_linear.", "rgb_linear."))}
    parts, floats = [], []
    pend = torch.zeros(160, device=dev)
    for l in range(2):
        p = f"transformer.layers.{l}."
        wqkv, wo = sd[p + "0.fn.fn.to_qkv.weight"], sd[p + "0.fn.fn.to_out.0.weight"]
        w1, w2 = sd[p + "1.fn.fn.net.0.weight"], sd[p + "1.fn.fn.net.3.weight"]
        # The LayerNorm affine is folded into the GEMM that consumes it: LN(x) W^T = xhat (W diag(g))^T + W b.  The
        # kernel's LayerNorm epilogue then writes only xhat = (x - mean) rstd (no gamma / beta loads and FMAs: a fifth
        # of its instructions) plus a constant 1.0 in the first pad column (K index 155), against which the q|k|v
        # weights carry W b as an extra K column; for the feed-forward the same column carries W b + b1, so the GELU
        # epilogue adds no bias either.
        g1, be1 = sd[p + "0.fn.norm.weight"], sd[p + "0.fn.norm.bias"]
        g2, be2 = sd[p + "1.fn.norm.weight"], sd[p + "1.fn.norm.bias"]
        wqkv = torch.cat([wqkv * g1[None, :], (wqkv @ be1)[:, None]], 1)          # (768, 156)
        b1_eff = sd[p + "1.fn.fn.net.0.bias"] + w1 @ be2
        w1 = torch.cat([w1 * g2[None, :], b1_eff[:, None]], 1)                      # (128, 156): the bias as K column 155
        qkv = [torch.cat([wqkv[64 * h:64 * h + 64], wqkv[256 + 64 * h:256 + 64 * h + 64],
                          wqkv[512 + 64 * h:512 + 64 * h + 64]], 0) for h in range(4)]
        out = [wo[:, 64 * h:64 * h + 64] for h in range(4)]
        # issue order of the attention block (two epilogue teams, csrc/dense_tc.cu): q|k|v of heads 0, 1, 2, then
        # Wo_0, q|k|v of head 3, Wo_1, Wo_2, Wo_3
        for item in ("q0", "q1", "q2", "o0", "q3", "o1", "o2", "o3"):
            h = int(item[1])
            # the out-projection runs as an fp16 x fp16 GEMM: its A operand, the attention output, is produced in fp16
            # by the epilogue and handed over as it is (kind::f16 does not mix an fp16 A with a bf16 B)
            parts.append(pack_kmajor_sw128(qkv[h], 192, 192) if item[0] == "q"
                         else pack_kmajor_sw128(out[h], 160, 64, dtype=torch.float16))
        parts.append(pack_kmajor_sw128(w1[:64], 64, 192))       # the FF hidden layer is computed in two 64-row halves
        parts.append(pack_kmajor_sw128(w1[64:128], 64, 192))
        parts.append(pack_kmajor_sw128(w2, 160, 128))
        pend_in = pend.clone()
        pend_mid = pend_in + _pad160(sd[p + "0.fn.fn.to_out.0.bias"])
        pend = pend_mid + _pad160(sd[p + "1.fn.fn.net.3.bias"])
        floats += [_pad160(sd[p + "0.fn.norm.weight"]), _pad160(sd[p + "0.fn.norm.bias"]), pend_in,
                   _pad160(sd[p + "1.fn.norm.weight"]), _pad160(sd[p + "1.fn.norm.bias"]), pend_mid,
                   b1_eff]          # (the ln_* vectors stay in the blob for reference; the kernels no longer read them)
    floats.append(pend)
    assert sum(x.numel() for x in parts) == T_BYTES

    def perm_x(w):      # columns [PE 39 | tok0 155] -> [tok0 155, 5 zeros, PE 39]
        o = torch.zeros(w.shape[0], 199, device=dev)
        o[:, :155] = w[:, 39:194]
        o[:, 160:199] = w[:, :39]
        return o

    W = [sd[f"pts_linears.{i}.weight"] for i in range(8)]

    def halves(*ws):      # per N-half (128 output rows): the K parts in issue order, 4 chunks of 64 each
        for h in range(2):
            for w in ws:
                parts.append(pack_kmajor_sw128(w[128 * h:128 * h + 128], 128, 256))

    halves(perm_x(W[0]))
    for i in range(1, 5):
        halves(W[i])
    halves(perm_x(W[5][:, :194]), W[5][:, 194:])
    halves(W[6])
    halves(W[7])
    halves(sd["feature_linear.weight"])
    wv = sd["views_linear.weight"]
    parts.append(pack_kmajor_sw128(wv[:, 256:411], 128, 256))
    parts.append(pack_kmajor_sw128(wv[:, :256], 128, 256))
    assert sum(x.numel() for x in parts) == T_BYTES + M_BYTES
    floats += [sd[f"pts_linears.{i}.bias"] for i in range(8)]
    floats += [sd["alpha_linear.weight"].reshape(-1), sd["feature_linear.bias"], sd["views_linear.bias"],
               sd["rgb_linear.weight"].reshape(-1), sd["alpha_linear.bias"].reshape(-1), sd["rgb_linear.bias"].reshape(-1)]
    f = torch.cat([x.reshape(-1).float() for x in floats])
    assert f.numel() == T_FLOATS + M_FLOATS, f.numel()
    blob = torch.cat(parts + [f.contiguous().view(torch.uint8)])
    assert blob.numel() == BLOB_BYTES
    return blob.contiguous()
