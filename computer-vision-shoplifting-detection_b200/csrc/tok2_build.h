// Host side of tokenizer v2: operand images built once per model + the per-(model, T) tile program.
#pragma once
#include <string>
#include <vector>

#include "sf_internal.h"
#include "tok2.h"

namespace sf {
namespace t2 {

struct BlockStatic {
  int cin, cout, cin_p, cp, stride, identity;
  uint32_t off_gcn, off_res;        // const blob: [cin_p/8][cp][8] K-major B images (block 0: unused)
  uint32_t off_bias_g, off_bias_o;  // const blob: fp32 [cp]
  uint32_t off_tcn, tcn_bytes;      // blob (after the const part): [cp/8][9 taps * cp][8], taps in descending order per stride phase
  int tap_pos[kTaps];               // block index of tap k inside the image
};

struct Static {
  bool ok = false;
  std::string why;
  int V = 0, WT = 0, rows = 0, c_in = 0, n_blocks = 0;
  BlockStatic blk[kMaxBlocks];
  uint32_t zero_lbo = 0;            // leading-dimension byte offset of the zero operand (0: both K halves read one plane)
  uint32_t off_ablk = 0, off_zero = 0, off_ell = 0, off_hc = 0, off_scale = 0, off_shift = 0;
  uint32_t off_g0tab_h = 0;         // fp16 format: the block-0 graph-conv table as halves, per 8 channels (w_x[8], w_y[8], b[8])
  uint32_t off_g0tab = 0, off_r0tab = 0;                            // block 0 (CUDA cores), fp32 [cp0 / 4][w_x[4], w_y[4], b[4]]
  int ell_width = 5;
  int pool_tokens = 0;              // > 0: AdaptiveAvgPool2d((pool_tokens, V)) after the last block
  bool f16 = true;                  // 16-bit operands (images and activations) are fp16, else bf16; see build_static
  uint32_t const_bytes = 0;
  std::vector<unsigned char> blob;  // const part [0, const_bytes) then the temporal-conv images
};

// `tok` must point at HOST copies of the folded fp32 weights.
void build_static(const Tokenizer& tok, int pool_tokens, bool allow_f16, Static* out);

struct Program {
  bool ok = false;
  std::string why;
  Plan plan;                        // const_src is filled in by the launcher
  std::vector<Mma> mma;
  std::vector<Group> groups;
  std::vector<Stage> stages[kTeams];
  std::vector<Load> loads;
};

void build_program(const Static& st, int T, int max_smem, Program* out);

// test infrastructure: executes the program on the host (bf16 operands, fp32 accumulate) for B windows.
// schedule_seed picks the interleaving of the three item sequences; returns false on deadlock / hazard.
bool emulate(const Static& st, const Program& pr, const float* poses, int64_t B, float* tokens, uint32_t schedule_seed,
             std::string* err);

}  // namespace t2
}  // namespace sf
