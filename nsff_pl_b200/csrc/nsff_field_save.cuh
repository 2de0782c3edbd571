// Layout of the activation buffer that the training-mode forward (nsff_field_forward_save, nsff_field_tc.cu) writes and
// the tensor-core backward (nsff_field_backward_tc, nsff_field_bwd_tc.cu) reads instead of recomputing the layers.
// Every plane is a bf16 tile image of the hidden activation ([128-point tile][4 chunks][128 x 128 B swizzled]) - the
// forward kernel's shared-memory hidden chunks ARE that image, so saving is a register -> global copy of what it writes
// to shared memory anyway - plus, for ReLU layers, the bit plane [tile][8 words][128 rows] of (activation > 0).
#pragma once
#include <stdint.h>
#include <stddef.h>

#include "nsff_b200.h"

namespace nsff {

constexpr size_t kSaveImgPerTile = 4 * 128 * 128;   // 4 chunks of 128 rows x 64 bf16
constexpr size_t kSaveBitsPerTile = 8 * 128 * 4;    // 8 words x 128 rows
constexpr uint64_t kNoSave = ~0ull;

struct SaveLayout {
  uint64_t s_h[8], s_hb[8], s_feat, s_dir, s_dirb;   // static trunk, final features, dir layer (+ bits)
  uint64_t t_h[8], t_hb[8], t_feat;                  // transient trunk, final features
  uint64_t total;
};

inline SaveLayout save_layout(const nsff_field_desc* d, int64_t B, int out_mask) {
  SaveLayout L;
  const uint64_t nt = (uint64_t)((B + 127) / 128);
  const uint64_t img = nt * kSaveImgPerTile, bits = nt * kSaveBitsPerTile;
  uint64_t o = 0;
  auto take = [&](bool on, uint64_t bytes) {
    if (!on) return kNoSave;
    const uint64_t at = o;
    o += bytes;
    return at;
  };
  const bool s = out_mask & NSFF_OUT_STATIC, t = out_mask & NSFF_OUT_TRANSIENT;
  for (int i = 0; i < 8; ++i) L.s_h[i] = take(s, img), L.s_hb[i] = take(s, bits);
  L.s_feat = take(s, img);
  L.s_dir = take(s && d->use_viewdir, img);
  L.s_dirb = take(s && d->use_viewdir, bits);
  for (int i = 0; i < 8; ++i) L.t_h[i] = take(t, img), L.t_hb[i] = take(t, bits);
  L.t_feat = take(t, img);
  L.total = o;
  return L;
}

}  // namespace nsff
