// Host-only: prints the per-tile program (sub-steps, K-chunk ops, explicit drain waits) that ProgramBuilder emits for the
// backward chain of a pass ("<static> <rows_pe> <rows_t>") or for a forward ("fwd <out_mask> <fuse> [save]"), without touching a GPU.  Build: nvcc -std=c++17 --expt-relaxed-constexpr
//   -gencode arch=compute_100a,code=sm_100a -Iinclude -Inspl_pl_b200/csrc -o tools/bin/dump_program tools/dump_program.cu
#include "../nsff_pl_b200/csrc/nsff_field_tc.cu"
namespace nsff {
thread_local cudaError_t g_last_cuda_error = cudaSuccess;
std::atomic<uint64_t> g_launch_count{0};
}
using namespace nsff::tc;

static PackLayer hid(std::initializer_list<uint8_t> chunks) {
  PackLayer pl; memset(&pl, 0, sizeof(pl));
  pl.present = true; pl.n = 256; pl.block_rows = 256; pl.blk_explicit = true;
  for (uint8_t c : chunks) { pl.chunk[pl.nchunks] = c; pl.blk[pl.nchunks] = (uint8_t)pl.nchunks; ++pl.nchunks; }
  return pl;
}
static PackLayer rows(int half) {
  PackLayer pl; memset(&pl, 0, sizeof(pl));
  pl.present = true; pl.n = 64; pl.block_rows = 128; pl.row0 = 64 * half; pl.blk_explicit = true; pl.nchunks = 4;
  for (int k = 0; k < 4; ++k) pl.chunk[k] = (uint8_t)(CH_H3 - k), pl.blk[k] = (uint8_t)(3 - k);
  return pl;
}
static void print_program(const FwdParams& P, int rc);

// "fwd <out_mask> <fuse> [save]": the forward program (inference: fuse = 1; training forward: fuse = 0, save = 1)
static int forward_main(int argc, char** argv) {
  const int mask = argc > 2 ? atoi(argv[2]) : 15, fuse = argc > 3 ? atoi(argv[3]) : 1, save = argc > 4 ? atoi(argv[4]) : 0;
  nsff_field_desc d;
  memset(&d, 0, sizeof(d));
  d.W = 256, d.D = 8, d.skip = 4, d.in_xyz = 63, d.in_dir = 27, d.in_a = 0, d.in_t = 48, d.use_viewdir = 1, d.has_transient = 1,
  d.has_flow = 1, d.flow_scale = 0.2f;
  PackLayout L;
  if (build_layout(d, L) != NSFF_OK) return 1;
  static FwdParams P;
  memset(&P, 0, sizeof(P));
  int ncols = 0;
  SaveLayout SL = save_layout(&d, 1024, mask);
  const int rc = build_program(d, L, mask, true, false, fuse != 0, P, &ncols, save ? &SL : nullptr);
  print_program(P, rc);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "fwd")) return forward_main(argc, argv);
  const bool want_s = argc > 1 && atoi(argv[1]), rows_pe = argc > 2 && atoi(argv[2]), rows_t = argc > 3 ? atoi(argv[3]) : 1;
  static FwdParams P; memset(&P, 0, sizeof(P));
  ProgramBuilder pb(P, true, false);
  const uint8_t DPS = CH_PE, DPT = CH_T;
  auto H4 = {(uint8_t)CH_H0, (uint8_t)CH_H1, (uint8_t)CH_H2, (uint8_t)CH_H3};
  if (want_s) {
    pb.push(hid({DPS}), EPI_MASK_HID, 0, 0);
    pb.push(hid(H4), EPI_LIN_HID, 0, 0);
    pb.push(hid({DPS, CH_H0, CH_H1, CH_H2, CH_H3}), EPI_MASK_HID, 0, 0);
    for (int i = 7; i >= 1; --i) pb.push(hid(H4), EPI_MASK_HID, 0, 0);
  }
  pb.push(hid({DPT}), EPI_LIN_HID, 0, 0);
  pb.push(hid(H4), EPI_MASK_HID, 0, 0);
  for (int i = 7; i >= 1; --i) {
    if (i == 4) {
      if (rows_pe) pb.push(rows(0), EPI_ROWS, kNoSave, kNoSave, 0);
      if (rows_t) pb.push(rows(1), EPI_ROWS, kNoSave, kNoSave, 1);
    }
    pb.push(hid(H4), EPI_MASK_HID, 0, 0);
  }
  if (rows_pe) pb.push(rows(0), EPI_ROWS, kNoSave, kNoSave, 0);
  if (rows_t) pb.push(rows(1), EPI_ROWS, kNoSave, kNoSave, 1);
  print_program(P, pb.finish());
  return 0;
}

static void print_program(const FwdParams& P, int rc) {
  printf("rc %d subs %d ops %d\n", rc, P.n_subs, P.n_ops);
  for (int i = 0; i < P.n_subs; ++i) {
    const SubStep& s = P.subs[i];
    printf("sub %2d acc %d d_col %3d epi %d n %3d arrive_empty %d\n", i, s.acc_id, s.d_col, s.epi, s.n, s.arrive_empty);
  }
  for (int i = 0; i < P.n_ops; ++i) {
    const ChunkOp& o = P.ops[i];
    printf("op %2d acc %d d_col %3d a %4d a2 %4d %s%s%s%s%s%s wait_hid %d wait_empty 0x%02x bytes %u\n", i, o.acc_id, o.d_col, o.a, o.a2,
           o.flags & OP_TS ? "TS " : "SS ", o.flags & OP_FIRST ? "FIRST " : "", o.flags & OP_LAST ? "LAST " : "",
           o.flags & OP_ADV_HID ? "ADV " : "", o.flags & OP_TWO ? "TWO " : "", o.flags & OP_EXTRA_DONE ? "XDONE " : "", o.wait_hid,
           o.wait_empty, o.bytes);
  }
}
