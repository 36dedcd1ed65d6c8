// MinAtar Breakout dynamics, one env per thread, state in registers.
//
// Restates gymnax==0.0.6 gymnax/environments/minatar/breakout.py (third party,
// not vendored in the reference; the reference only calls it through
// gymnax.make at purejaxql/pqn_minatar.py:103 and vmap_step/vmap_reset at
// :107-112).  Integer state, integer arithmetic: results are bit-exact.
//
// State words (SoA, see env_common.cuh):
//   w0  ball_y[0:4) ball_x[4:8) ball_dir[8:10) pos[10:14) last_y[14:18)
//       last_x[18:22) strike[22] terminal[23]
//   w1  time (int32)
//   w2..w5  brick_map as 100 bits, bit p = y*10 + x (gymnax keeps f32[10,10])
// Observation (10,10,4) float32 in gymnax; here produced as 400 bits, bit index
// (y*10+x)*4 + c  == flat index of the (10,10,4) array, c: 0 paddle, 1 ball,
// 2 trail, 3 brick.
#pragma once
#include "env_common.cuh"

namespace pqn {

struct BreakoutEnv {
  static constexpr int ID = ENV_BREAKOUT;
  static constexpr int CORE_WORDS = 6;
  static constexpr int STATE_WORDS = CORE_WORDS + LOG_WORDS;
  static constexpr int NUM_ACTIONS = 3;  // minimal action set [0,1,3] = noop,left,right
  static constexpr int OBS_H = 10, OBS_W = 10, OBS_C = 4;
  static constexpr int OBS_DIM = 400;
  static constexpr bool BINARY_OBS = true;
  static constexpr bool OBS_IN_REGS = true;  // obs_bits() fills a register array (fixed cell count)
  static constexpr int OBS_WORDS = 13;      // ceil(400/32)
  static constexpr int OBS_WORDS_PAD = 16;  // 64-byte rows in the packed rollout buffer
  static constexpr int DEFAULT_MAX_STEPS = 1000;

  struct State {
    int ball_y, ball_x, ball_dir, pos, last_y, last_x;
    bool strike, terminal;
    int time;
    uint32_t brick[4];
  };

  template <typename W>
  PQN_HD static void load(State& s, const W* __restrict__ st, int64_t N, int64_t i) {
    const uint32_t w = st[i];
    s.ball_y = w & 15u; s.ball_x = (w >> 4) & 15u; s.ball_dir = (w >> 8) & 3u;
    s.pos = (w >> 10) & 15u; s.last_y = (w >> 14) & 15u; s.last_x = (w >> 18) & 15u;
    s.strike = (w >> 22) & 1u; s.terminal = (w >> 23) & 1u;
    s.time = (int)st[N + i];
#pragma unroll
    for (int k = 0; k < 4; ++k) s.brick[k] = st[(int64_t)(2 + k) * N + i];
  }
  PQN_HD static void store(const State& s, uint32_t* __restrict__ st, int64_t N, int64_t i) {
    st[i] = (uint32_t)s.ball_y | ((uint32_t)s.ball_x << 4) | ((uint32_t)s.ball_dir << 8) |
            ((uint32_t)s.pos << 10) | ((uint32_t)s.last_y << 14) | ((uint32_t)s.last_x << 18) |
            ((uint32_t)s.strike << 22) | ((uint32_t)s.terminal << 23);
    st[N + i] = (uint32_t)s.time;
#pragma unroll
    for (int k = 0; k < 4; ++k) st[(int64_t)(2 + k) * N + i] = s.brick[k];
  }

  PQN_HD static void fill_bricks(State& s) {  // brick_map.at[1:4, :].set(1)
    s.brick[0] |= 0xFFFFFC00u;                // bits 10..31
    s.brick[1] |= 0x000000FFu;                // bits 32..39
  }

  // reset_env: ball_start = jax.random.choice(key, [0,1]) == randint(key,(),0,2)
  PQN_HD static void reset_env(Key key, int part, int /*max_steps*/, State& s) {
    const int start = randint_scalar(key, 2u, part);
    s.ball_y = 3; s.ball_x = start ? 9 : 0; s.ball_dir = start ? 3 : 2; s.pos = 4;
    s.brick[0] = s.brick[1] = s.brick[2] = s.brick[3] = 0u;
    fill_bricks(s);
    s.strike = false; s.last_y = 3; s.last_x = s.ball_x; s.time = 0; s.terminal = false;
  }

  // step_env = step_agent + step_ball_brick + time/terminal bookkeeping.
  PQN_HD static void step_env(Key /*key*/, int /*part*/, int max_steps, State& s, int action,
                              float& reward, bool& done) {
    const int a = action <= 0 ? 0 : (action == 1 ? 1 : 3);
    // ---- step_agent
    int pos = s.pos;
    if (a == 1) pos = pos - 1 < 0 ? 0 : pos - 1;
    else if (a == 3) pos = pos + 1 > 9 ? 9 : pos + 1;
    const int last_x = s.ball_x, last_y = s.ball_y;
    int dir = s.ball_dir;
    int new_x = (dir == 0 || dir == 3) ? s.ball_x - 1 : s.ball_x + 1;
    int new_y = (dir == 0 || dir == 1) ? s.ball_y - 1 : s.ball_y + 1;
    const bool border_x = new_x < 0 || new_x > 9;
    if (border_x) { new_x = new_x < 0 ? 0 : 9; dir ^= 1; }          // [1,0,3,2][dir]
    // ---- step_ball_brick
    const bool border_y = new_y < 0;
    if (border_y) { new_y = 0; dir = 3 - dir; }                     // [3,2,1,0][dir]
    const int yi = new_y > 9 ? 9 : new_y;                           // XLA gather clamps
    const int p = yi * 10 + new_x;
    const bool brick_here = (s.brick[p >> 5] >> (p & 31)) & 1u;
    const bool strike_toggle = !border_y && brick_here;
    const bool strike_bool = !s.strike && strike_toggle;
    reward = strike_bool ? 1.0f : 0.0f;
    if (strike_bool) {
      s.brick[p >> 5] &= ~(1u << (p & 31));
      new_y = last_y;
      dir = 3 - dir;
    }
    const bool brick_cond = !strike_toggle && new_y == 9;
    if (brick_cond && (s.brick[0] | s.brick[1] | s.brick[2] | s.brick[3]) == 0u) fill_bricks(s);
    const bool redirect1 = brick_cond && s.ball_x == pos;           // old ball_x, new pos
    if (redirect1) { dir = 3 - dir; new_y = last_y; }
    const bool redirect2 = brick_cond && !redirect1 && new_x == pos;
    if (redirect2) { dir ^= 2; new_y = last_y; }                    // [2,3,0,1][dir]
    const bool terminal = brick_cond && !redirect1 && !redirect2;
    s.pos = pos; s.last_x = last_x; s.last_y = last_y; s.ball_dir = dir;
    s.strike = strike_toggle; s.ball_x = new_x; s.ball_y = new_y;
    s.time = s.time + 1;
    done = terminal || s.time >= max_steps;
    s.terminal = done;
  }

  // 8 consecutive bits -> one bit per nibble (bit i -> bit 4i).
  PQN_HD static uint32_t spread8(uint32_t b) {
    uint32_t x = b & 0xFFu;
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
  }

  // get_obs as packed bits; o[w] covers pixels 8w..8w+7, 4 channel bits each.
  PQN_HD static void obs_bits(const State& s, uint32_t (&o)[OBS_WORDS_PAD]) {
    const int p_pad = 90 + s.pos;
    const int p_ball = s.ball_y * 10 + s.ball_x;
    const int p_trail = s.last_y * 10 + s.last_x;
#pragma unroll
    for (int w = 0; w < OBS_WORDS_PAD; ++w) {
      if (w >= OBS_WORDS) { o[w] = 0u; continue; }
      uint32_t v = spread8(s.brick[w >> 2] >> (8 * (w & 3))) << 3;
      if ((p_pad >> 3) == w) v |= 1u << (4 * (p_pad & 7) + 0);
      if ((p_ball >> 3) == w) v |= 1u << (4 * (p_ball & 7) + 1);
      if ((p_trail >> 3) == w) v |= 1u << (4 * (p_trail & 7) + 2);
      o[w] = v;
    }
  }
};

}  // namespace pqn
