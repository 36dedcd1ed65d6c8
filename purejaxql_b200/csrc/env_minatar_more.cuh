// MinAtar Freeway / SpaceInvaders / Asterix dynamics, one env per thread.
//
// Restated from gymnax==0.0.6 gymnax/environments/minatar/{freeway,space_invaders,asterix}.py (third party, JAX
// ports of the MinAtar games; reached by the reference only through gymnax.make, purejaxql/pqn_minatar.py:103).
// Integer state and arithmetic.  PARITY UNPINNED against a live gymnax (not installable here): the tests pin
// these kernels bit-exactly to the CPU checker under tests (a NumPy restatement of the same published algorithms).
//
// Observations are produced as bits into a caller-provided word array (`o[w * stride]`, bit index = flat index
// of the (10,10,C) observation), because these games set a data-dependent number of cells.
#pragma once
#include "env_common.cuh"

namespace pqn {

PQN_HD void obs_set_bit(uint32_t* o, int stride, int idx) { o[(idx >> 5) * stride] |= 1u << (idx & 31); }

// jax.random.randint(key, (n,), lo, hi)[i] with the key split hoisted
struct RandintKeys {
  Key k1, k2;
};
PQN_HD RandintKeys randint_keys(Key k, int part) {
  RandintKeys r;
  split2(k, part, r.k1, r.k2);
  return r;
}
PQN_HD int32_t randint_elem(const RandintKeys& rk, uint32_t n, uint32_t i, int32_t lo, int32_t hi, int part) {
  const uint32_t hb = bits_at(rk.k1, n, i, part), lb = bits_at(rk.k2, n, i, part);
  const uint32_t span = (uint32_t)(hi - lo);
  uint32_t mult = 65536u % span;
  mult = (mult * mult) % span;
  return lo + (int32_t)(((hb % span) * mult + (lb % span)) % span);
}

// ---------------------------------------------------------------------------
// Freeway: state words  w0 pos[0:4) move_timer[4:6) terminal[6] ; w1 time ;
//          w2..w5 two cars per word (16 bits each): x[0:4) timer[4:7) speed+5[7:11)   (car c sits in row y = c+1)
// ---------------------------------------------------------------------------
struct FreewayEnv {
  static constexpr int ID = ENV_FREEWAY;
  static constexpr int CORE_WORDS = 6;
  static constexpr int STATE_WORDS = CORE_WORDS + LOG_WORDS;
  static constexpr int NUM_ACTIONS = 3;  // minimal action set [0,2,4] = noop, up, down
  static constexpr int OBS_H = 10, OBS_W = 10, OBS_C = 7;
  static constexpr int OBS_DIM = 700;
  static constexpr bool BINARY_OBS = true;
  static constexpr bool OBS_IN_REGS = false;
  static constexpr int OBS_WORDS = 22;
  static constexpr int OBS_WORDS_PAD = 24;
  static constexpr int DEFAULT_MAX_STEPS = 2500;
  static constexpr int PLAYER_SPEED = 3;

  struct State {
    int pos, move_timer, time;
    bool terminal;
    int cx[8], ct[8], cs[8];
  };

  template <typename W>
  PQN_HD static void load(State& s, const W* __restrict__ st, int64_t N, int64_t i) {
    const uint32_t w = st[i];
    s.pos = w & 15u; s.move_timer = (w >> 4) & 3u; s.terminal = (w >> 6) & 1u;
    s.time = (int)st[N + i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t v = st[(int64_t)(2 + k) * N + i];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t c = (v >> (16 * h)) & 0xFFFFu;
        s.cx[2 * k + h] = c & 15u; s.ct[2 * k + h] = (c >> 4) & 7u; s.cs[2 * k + h] = (int)((c >> 7) & 15u) - 5;
      }
    }
  }
  PQN_HD static void store(const State& s, uint32_t* __restrict__ st, int64_t N, int64_t i) {
    st[i] = (uint32_t)s.pos | ((uint32_t)s.move_timer << 4) | ((uint32_t)s.terminal << 6);
    st[N + i] = (uint32_t)s.time;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t v = 0u;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = 2 * k + h;
        v |= ((uint32_t)s.cx[c] | ((uint32_t)s.ct[c] << 4) | ((uint32_t)(s.cs[c] + 5) << 7)) << (16 * h);
      }
      st[(int64_t)(2 + k) * N + i] = v;
    }
  }

  // speeds = randint(key_speed,(8,),1,6) * choice(key_dirs,[-1,1],(8,)) with key_speed,key_dirs = split(key)
  PQN_HD static void draw_speeds(Key key, int part, int (&sp)[8]) {
    Key ks, kd;
    split2(key, part, ks, kd);
    const RandintKeys rs = randint_keys(ks, part), rd = randint_keys(kd, part);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sp[c] = randint_elem(rs, 8u, c, 1, 6, part) * (randint_elem(rd, 8u, c, 0, 2, part) * 2 - 1);
  }

  PQN_HD static void reset_env(Key key, int part, int /*max_steps*/, State& s) {
    int sp[8];
    draw_speeds(key, part, sp);
#pragma unroll
    for (int c = 0; c < 8; ++c) { s.cx[c] = 0; s.ct[c] = sp[c] < 0 ? -sp[c] : sp[c]; s.cs[c] = sp[c]; }
    s.pos = 9; s.move_timer = PLAYER_SPEED; s.time = 0; s.terminal = false;
  }

  PQN_HD static void step_env(Key key, int part, int max_steps, State& s, int action, float& reward, bool& done) {
    const int a = action <= 0 ? 0 : (action == 1 ? 2 : 4);
    const bool up = a == 2 && s.move_timer == 0, down = a == 4 && s.move_timer == 0;
    if (up || down) s.move_timer = PLAYER_SPEED;
    if (up) s.pos = s.pos - 1 < 0 ? 0 : s.pos - 1;
    else if (down) s.pos = s.pos + 1 > 9 ? 9 : s.pos + 1;
    const bool win = s.pos == 0;
    reward = win ? 1.0f : 0.0f;
    if (win) {
      s.pos = 9;
      int sp[8];
      draw_speeds(key, part, sp);  // gymnax samples every step and selects on win: same values
#pragma unroll
      for (int c = 0; c < 8; ++c) { s.ct[c] = sp[c] < 0 ? -sp[c] : sp[c]; s.cs[c] = sp[c]; }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (s.cx[c] == 4 && c + 1 == s.pos) s.pos = 9;
      if (s.ct[c] == 0) {
        s.ct[c] = s.cs[c] < 0 ? -s.cs[c] : s.cs[c];
        int nx = s.cx[c] + (s.cs[c] > 0 ? 1 : -1);
        nx = nx < 0 ? 9 : (nx > 9 ? 0 : nx);
        s.cx[c] = nx;
        if (s.cx[c] == 4 && c + 1 == s.pos) s.pos = 9;
      } else {
        s.ct[c] -= 1;
      }
    }
    s.move_timer -= s.move_timer > 0 ? 1 : 0;
    s.time += 1;
    done = s.time >= max_steps;
    s.terminal = done;
  }

  PQN_HD static void obs_bits_mem(const State& s, uint32_t* o, int stride) {
    obs_set_bit(o, stride, (s.pos * 10 + 4) * OBS_C + 0);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int y = c + 1, x = s.cx[c];
      obs_set_bit(o, stride, (y * 10 + x) * OBS_C + 1);
      int back = s.cs[c] > 0 ? x - 1 : x + 1;
      back = back < 0 ? 9 : (back > 9 ? 0 : back);
      const int sp = s.cs[c] < 0 ? -s.cs[c] : s.cs[c];
      obs_set_bit(o, stride, (y * 10 + back) * OBS_C + 1 + sp);
    }
  }
};

// ---------------------------------------------------------------------------
// 100-bit board maps (bit p = y*10 + x) held in 4 words
// ---------------------------------------------------------------------------
struct Map100 {
  uint32_t w[4];
};
PQN_HD uint32_t m_word(const Map100& m, int k) { return k == 0 ? m.w[0] : (k == 1 ? m.w[1] : (k == 2 ? m.w[2] : m.w[3])); }
PQN_HD bool m_get(const Map100& m, int p) { return (m_word(m, p >> 5) >> (p & 31)) & 1u; }
PQN_HD void m_set(Map100& m, int p) {
  const uint32_t b = 1u << (p & 31);
  const int k = p >> 5;
  m.w[0] |= k == 0 ? b : 0u; m.w[1] |= k == 1 ? b : 0u; m.w[2] |= k == 2 ? b : 0u; m.w[3] |= k == 3 ? b : 0u;
}
PQN_HD bool m_any(const Map100& m) { return (m.w[0] | m.w[1] | m.w[2] | m.w[3]) != 0u; }
PQN_HD int m_count(const Map100& m) {
#if defined(__CUDA_ARCH__)
  return __popc(m.w[0]) + __popc(m.w[1]) + __popc(m.w[2]) + __popc(m.w[3]);
#else
  return __builtin_popcount(m.w[0]) + __builtin_popcount(m.w[1]) + __builtin_popcount(m.w[2]) + __builtin_popcount(m.w[3]);
#endif
}
PQN_HD Map100 m_and(const Map100& a, const Map100& b) { return Map100{{a.w[0] & b.w[0], a.w[1] & b.w[1], a.w[2] & b.w[2], a.w[3] & b.w[3]}}; }
PQN_HD Map100 m_andnot(const Map100& a, const Map100& b) { return Map100{{a.w[0] & ~b.w[0], a.w[1] & ~b.w[1], a.w[2] & ~b.w[2], a.w[3] & ~b.w[3]}}; }
PQN_HD Map100 m_or(const Map100& a, const Map100& b) { return Map100{{a.w[0] | b.w[0], a.w[1] | b.w[1], a.w[2] | b.w[2], a.w[3] | b.w[3]}}; }
PQN_HD Map100 m_mask100(Map100 m) { m.w[3] &= 0xFu; return m; }  // bits 96..99 only
PQN_HD Map100 m_shr(const Map100& m, int n) {  // 0 < n < 32
  return Map100{{(m.w[0] >> n) | (m.w[1] << (32 - n)), (m.w[1] >> n) | (m.w[2] << (32 - n)),
                 (m.w[2] >> n) | (m.w[3] << (32 - n)), m.w[3] >> n}};
}
PQN_HD Map100 m_shl(const Map100& m, int n) {  // 0 < n < 32, result masked to 100 bits
  return m_mask100(Map100{{m.w[0] << n, (m.w[1] << n) | (m.w[0] >> (32 - n)), (m.w[2] << n) | (m.w[1] >> (32 - n)),
                           (m.w[3] << n) | (m.w[2] >> (32 - n))}});
}
// column x of every row: bits x, x+10, ..., x+90
PQN_HD Map100 m_col(int x) {
  Map100 c{{0u, 0u, 0u, 0u}};
#pragma unroll
  for (int y = 0; y < 10; ++y) m_set(c, y * 10 + x);
  return c;
}
PQN_HD Map100 m_row(int y) {
  Map100 r{{0u, 0u, 0u, 0u}};
#pragma unroll
  for (int x = 0; x < 10; ++x) m_set(r, y * 10 + x);
  return r;
}

// ---------------------------------------------------------------------------
// SpaceInvaders: deterministic.  state words: w0 pos[0:4) dir_pos[4] enemy_move_interval[5:9)
//   alien_move_timer[9:13) alien_shot_timer[13:17) shot_timer[17:20) terminal[20] ramp_index[21:25) ; w1 time ;
//   w2..5 alien map, w6..9 friendly bullets, w10..13 enemy bullets
// ---------------------------------------------------------------------------
struct SpaceInvadersEnv {
  static constexpr int ID = ENV_SPACE_INVADERS;
  static constexpr int CORE_WORDS = 14;
  static constexpr int STATE_WORDS = CORE_WORDS + LOG_WORDS;
  static constexpr int NUM_ACTIONS = 4;  // minimal action set [0,1,3,5] = noop, left, right, fire
  static constexpr int OBS_H = 10, OBS_W = 10, OBS_C = 6;
  static constexpr int OBS_DIM = 600;
  static constexpr bool BINARY_OBS = true;
  static constexpr bool OBS_IN_REGS = false;
  static constexpr int OBS_WORDS = 19;
  static constexpr int OBS_WORDS_PAD = 20;
  static constexpr int DEFAULT_MAX_STEPS = 1000;
  static constexpr int SHOT_COOL_DOWN = 5, ENEMY_MOVE_INTERVAL = 12, ENEMY_SHOT_INTERVAL = 10;

  struct State {
    int pos, alien_dir, enemy_move_interval, alien_move_timer, alien_shot_timer, shot_timer, ramp_index, time;
    bool terminal;
    Map100 alien, fb, eb;
  };

  template <typename W>
  PQN_HD static void load(State& s, const W* __restrict__ st, int64_t N, int64_t i) {
    const uint32_t w = st[i];
    s.pos = w & 15u; s.alien_dir = ((w >> 4) & 1u) ? 1 : -1; s.enemy_move_interval = (w >> 5) & 15u;
    s.alien_move_timer = (w >> 9) & 15u; s.alien_shot_timer = (w >> 13) & 15u; s.shot_timer = (w >> 17) & 7u;
    s.terminal = (w >> 20) & 1u; s.ramp_index = (w >> 21) & 15u;
    s.time = (int)st[N + i];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s.alien.w[k] = st[(int64_t)(2 + k) * N + i];
      s.fb.w[k] = st[(int64_t)(6 + k) * N + i];
      s.eb.w[k] = st[(int64_t)(10 + k) * N + i];
    }
  }
  PQN_HD static void store(const State& s, uint32_t* __restrict__ st, int64_t N, int64_t i) {
    st[i] = (uint32_t)s.pos | ((s.alien_dir > 0 ? 1u : 0u) << 4) | ((uint32_t)s.enemy_move_interval << 5) |
            ((uint32_t)s.alien_move_timer << 9) | ((uint32_t)s.alien_shot_timer << 13) | ((uint32_t)s.shot_timer << 17) |
            ((uint32_t)s.terminal << 20) | ((uint32_t)s.ramp_index << 21);
    st[N + i] = (uint32_t)s.time;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      st[(int64_t)(2 + k) * N + i] = s.alien.w[k];
      st[(int64_t)(6 + k) * N + i] = s.fb.w[k];
      st[(int64_t)(10 + k) * N + i] = s.eb.w[k];
    }
  }

  PQN_HD static void spawn_aliens(Map100& a) {  // alien_map[0:4, 2:8] = 1
#pragma unroll
    for (int y = 0; y < 4; ++y)
#pragma unroll
      for (int x = 2; x < 8; ++x) m_set(a, y * 10 + x);
  }

  PQN_HD static void reset_env(Key /*key*/, int /*part*/, int /*max_steps*/, State& s) {
    s.pos = 5; s.alien_dir = -1; s.enemy_move_interval = ENEMY_MOVE_INTERVAL; s.alien_move_timer = ENEMY_MOVE_INTERVAL;
    s.alien_shot_timer = ENEMY_SHOT_INTERVAL; s.shot_timer = 0; s.ramp_index = 0; s.time = 0; s.terminal = false;
    s.alien = Map100{{0u, 0u, 0u, 0u}}; s.fb = s.alien; s.eb = s.alien;
    spawn_aliens(s.alien);
  }

  PQN_HD static void step_env(Key /*key*/, int /*part*/, int max_steps, State& s, int action, float& reward, bool& done) {
    const int a = action <= 0 ? 0 : (action == 1 ? 1 : (action == 2 ? 3 : 5));
    // ---- player
    const bool fire = a == 5 && s.shot_timer == 0;
    if (fire) { m_set(s.fb, 90 + s.pos); s.shot_timer = SHOT_COOL_DOWN; }
    else if (a == 1) s.pos = s.pos - 1 < 0 ? 0 : s.pos - 1;
    else if (a == 3) s.pos = s.pos + 1 > 9 ? 9 : s.pos + 1;
    // ---- bullets: friendly up one row, enemy down one row
    s.fb = m_shr(s.fb, 10);
    s.eb = m_shl(s.eb, 10);
    bool terminal = m_get(s.eb, 90 + s.pos);
    // ---- aliens
    terminal = terminal || m_get(s.alien, 90 + s.pos);
    const bool move = s.alien_move_timer == 0;
    if (move) {
      const int cnt = m_count(s.alien);
      s.alien_move_timer = cnt < s.enemy_move_interval ? cnt : s.enemy_move_interval;
      const bool edge = (m_any(m_and(s.alien, m_col(0))) && s.alien_dir < 0) ||
                        (m_any(m_and(s.alien, m_col(9))) && s.alien_dir > 0);
      if (edge) {
        s.alien_dir = -s.alien_dir;
        if (m_any(m_and(s.alien, m_row(9)))) terminal = true;
        // np.roll(alien_map, 1, axis=0): one row down, row 9 wraps to row 0
        s.alien = m_or(m_shl(s.alien, 10), m_shr(m_shr(m_shr(s.alien, 30), 30), 30));
      } else if (s.alien_dir < 0) {
        // np.roll(alien_map, -1, axis=1): one column left, column 0 wraps to column 9
        const Map100 c0 = m_and(s.alien, m_col(0));
        s.alien = m_or(m_andnot(m_shr(s.alien, 1), m_col(9)), m_shl(c0, 9));
      } else {
        const Map100 c9 = m_and(s.alien, m_col(9));
        s.alien = m_or(m_andnot(m_shl(s.alien, 1), m_col(0)), m_shr(c9, 9));
      }
      terminal = terminal || m_get(s.alien, 90 + s.pos);
    }
    // ---- alien shot from the alien nearest to the cannon (ties: lower column first)
    if (s.alien_shot_timer == 0) {
      s.alien_shot_timer = ENEMY_SHOT_INTERVAL;
      int col = -1;
      for (int d = 0; d < 10 && col < 0; ++d) {
        const int cl = s.pos - d, cr = s.pos + d;
        if (cl >= 0 && m_any(m_and(s.alien, m_col(cl)))) col = cl;
        else if (d > 0 && cr <= 9 && m_any(m_and(s.alien, m_col(cr)))) col = cr;
      }
      if (col >= 0) {
        int yy = 9;
        while (yy > 0 && !m_get(s.alien, yy * 10 + col)) --yy;
        m_set(s.eb, yy * 10 + col);
      }
    }
    // ---- kills
    const Map100 kill = m_and(s.alien, s.fb);
    reward = (float)m_count(kill);
    s.alien = m_andnot(s.alien, kill);
    s.fb = m_andnot(s.fb, kill);
    // ---- timers, ramping, respawn
    s.shot_timer -= s.shot_timer > 0 ? 1 : 0;
    s.alien_move_timer -= 1;
    s.alien_shot_timer -= 1;
    if (!m_any(s.alien)) {
      if (s.enemy_move_interval > 6) { s.enemy_move_interval -= 1; s.ramp_index += 1; }
      spawn_aliens(s.alien);
    }
    s.time += 1;
    done = terminal || s.time >= max_steps;
    s.terminal = done;
  }

  PQN_HD static void obs_bits_mem(const State& s, uint32_t* o, int stride) {
    obs_set_bit(o, stride, (90 + s.pos) * OBS_C + 0);
    const int dir_ch = s.alien_dir < 0 ? 2 : 3;
    for (int k = 0; k < 4; ++k) {
      uint32_t a = s.alien.w[k], f = s.fb.w[k], e = s.eb.w[k];
      while (a) {
#if defined(__CUDA_ARCH__)
        const int b = __ffs(a) - 1;
#else
        const int b = __builtin_ffs(a) - 1;
#endif
        a &= a - 1u;
        const int p = k * 32 + b;
        obs_set_bit(o, stride, p * OBS_C + 1);
        obs_set_bit(o, stride, p * OBS_C + dir_ch);
      }
      while (f) {
#if defined(__CUDA_ARCH__)
        const int b = __ffs(f) - 1;
#else
        const int b = __builtin_ffs(f) - 1;
#endif
        f &= f - 1u;
        obs_set_bit(o, stride, (k * 32 + b) * OBS_C + 4);
      }
      while (e) {
#if defined(__CUDA_ARCH__)
        const int b = __ffs(e) - 1;
#else
        const int b = __builtin_ffs(e) - 1;
#endif
        e &= e - 1u;
        obs_set_bit(o, stride, (k * 32 + b) * OBS_C + 5);
      }
    }
  }
};

// ---------------------------------------------------------------------------
// Asterix.  state words: w0 player_x[0:4) player_y[4:8) spawn_speed[8:12) spawn_timer[12:16) move_speed[16:19)
//   move_timer[19:22) shot_timer[22:25) terminal[25] ; w1 ramp_timer+1 [0:8) ramp_index[8:16) ; w2 time ;
//   w3,w4 entities, 8 bits each: x[0:4) lr[4] gold[5] filled[6]   (entity e lives in row y = e+1)
// RNG per step: key_lr, key_gold, key_slot = split(key, 3); lr = choice([1,0]); is_gold = choice([1,0], p=[1/3,2/3]);
// slot = choice(8, p = free/sum(free))  (jax.random.choice with p: cumsum + searchsorted on total*(1-uniform)).
// ---------------------------------------------------------------------------
struct AsterixEnv {
  static constexpr int ID = ENV_ASTERIX;
  static constexpr int CORE_WORDS = 5;
  static constexpr int STATE_WORDS = CORE_WORDS + LOG_WORDS;
  static constexpr int NUM_ACTIONS = 5;  // noop, left, up, right, down
  static constexpr int OBS_H = 10, OBS_W = 10, OBS_C = 4;
  static constexpr int OBS_DIM = 400;
  static constexpr bool BINARY_OBS = true;
  static constexpr bool OBS_IN_REGS = false;
  static constexpr int OBS_WORDS = 13;
  static constexpr int OBS_WORDS_PAD = 16;
  static constexpr int DEFAULT_MAX_STEPS = 1000;
  static constexpr int RAMP_INTERVAL = 100, INIT_SPAWN_SPEED = 10, INIT_MOVE_INTERVAL = 5;

  struct State {
    int player_x, player_y, shot_timer, spawn_speed, spawn_timer, move_speed, move_timer, ramp_timer, ramp_index, time;
    bool terminal;
    int ex[8];
    bool elr[8], egold[8], efill[8];
  };

  template <typename W>
  PQN_HD static void load(State& s, const W* __restrict__ st, int64_t N, int64_t i) {
    const uint32_t w = st[i], w1 = st[N + i];
    s.player_x = w & 15u; s.player_y = (w >> 4) & 15u; s.spawn_speed = (w >> 8) & 15u; s.spawn_timer = (w >> 12) & 15u;
    s.move_speed = (w >> 16) & 7u; s.move_timer = (w >> 19) & 7u; s.shot_timer = (w >> 22) & 7u; s.terminal = (w >> 25) & 1u;
    s.ramp_timer = (int)(w1 & 255u) - 1; s.ramp_index = (w1 >> 8) & 255u;
    s.time = (int)st[2 * N + i];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint32_t v = st[(int64_t)(3 + k) * N + i];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const uint32_t e = (v >> (8 * h)) & 255u;
        s.ex[4 * k + h] = e & 15u; s.elr[4 * k + h] = (e >> 4) & 1u; s.egold[4 * k + h] = (e >> 5) & 1u;
        s.efill[4 * k + h] = (e >> 6) & 1u;
      }
    }
  }
  PQN_HD static void store(const State& s, uint32_t* __restrict__ st, int64_t N, int64_t i) {
    st[i] = (uint32_t)s.player_x | ((uint32_t)s.player_y << 4) | ((uint32_t)s.spawn_speed << 8) |
            ((uint32_t)s.spawn_timer << 12) | ((uint32_t)s.move_speed << 16) | ((uint32_t)s.move_timer << 19) |
            ((uint32_t)s.shot_timer << 22) | ((uint32_t)s.terminal << 25);
    st[N + i] = (uint32_t)(s.ramp_timer + 1) | ((uint32_t)s.ramp_index << 8);
    st[2 * N + i] = (uint32_t)s.time;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      uint32_t v = 0u;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int e = 4 * k + h;
        // emptied slots are all-zero (gymnax zeroes the whole entity row)
        const uint32_t b = s.efill[e] ? ((uint32_t)s.ex[e] | ((uint32_t)s.elr[e] << 4) | ((uint32_t)s.egold[e] << 5) | (1u << 6)) : 0u;
        v |= b << (8 * h);
      }
      st[(int64_t)(3 + k) * N + i] = v;
    }
  }

  PQN_HD static void reset_env(Key /*key*/, int /*part*/, int /*max_steps*/, State& s) {
    s.player_x = 5; s.player_y = 5; s.shot_timer = 0; s.spawn_speed = INIT_SPAWN_SPEED; s.spawn_timer = INIT_SPAWN_SPEED;
    s.move_speed = INIT_MOVE_INTERVAL; s.move_timer = INIT_MOVE_INTERVAL; s.ramp_timer = RAMP_INTERVAL; s.ramp_index = 0;
    s.time = 0; s.terminal = false;
#pragma unroll
    for (int e = 0; e < 8; ++e) { s.ex[e] = 0; s.elr[e] = false; s.egold[e] = false; s.efill[e] = false; }
  }

  PQN_HD static void collide(State& s, int e, float& reward, bool& terminal) {
    if (s.efill[e] && s.ex[e] == s.player_x && e + 1 == s.player_y) {
      if (s.egold[e]) { reward += 1.0f; s.efill[e] = false; s.ex[e] = 0; s.elr[e] = false; s.egold[e] = false; }
      else terminal = true;
    }
  }

  PQN_HD static void step_env(Key key, int part, int max_steps, State& s, int action, float& reward, bool& done) {
    // ---- spawn
    if (s.spawn_timer == 0) {
      int nfree = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) nfree += s.efill[e] ? 0 : 1;
      if (nfree > 0) {
        Key kl, kg, ks;
        split3(key, part, kl, kg, ks);
        const int lr = 1 - randint_scalar(kl, 2u, part);
        const float c0 = (float)(1.0 / 3.0), c1 = c0 + (float)(2.0 / 3.0);
        const float rg = c1 * (1.0f - uniform_scalar(kg, part));
        const int is_gold = 1 - ((c0 < rg ? 1 : 0) + (c1 < rg ? 1 : 0));
        const float inv = 1.0f / (float)nfree;
        float cum = 0.f, cums[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { cum = cum + (s.efill[e] ? 0.0f : inv); cums[e] = cum; }
        const float rs = cum * (1.0f - uniform_scalar(ks, part));
        int slot = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) slot += cums[e] < rs ? 1 : 0;
        slot = slot > 7 ? 7 : slot;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (e == slot) { s.ex[e] = lr ? 0 : 9; s.elr[e] = lr != 0; s.egold[e] = is_gold != 0; s.efill[e] = true; }
      }
      s.spawn_timer = s.spawn_speed;
    }
    // ---- player
    if (action == 1) s.player_x = s.player_x - 1 < 0 ? 0 : s.player_x - 1;
    else if (action == 3) s.player_x = s.player_x + 1 > 9 ? 9 : s.player_x + 1;
    else if (action == 2) s.player_y = s.player_y - 1 < 1 ? 1 : s.player_y - 1;
    else if (action == 4) s.player_y = s.player_y + 1 > 8 ? 8 : s.player_y + 1;
    reward = 0.f;
    bool terminal = false;
#pragma unroll
    for (int e = 0; e < 8; ++e) collide(s, e, reward, terminal);
    if (s.move_timer == 0) {
      s.move_timer = s.move_speed;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (s.efill[e]) {
          s.ex[e] += s.elr[e] ? 1 : -1;
          if (s.ex[e] < 0 || s.ex[e] > 9) { s.efill[e] = false; s.ex[e] = 0; s.elr[e] = false; s.egold[e] = false; }
        }
        collide(s, e, reward, terminal);
      }
    }
    s.spawn_timer -= 1;
    s.move_timer -= 1;
    // ---- difficulty ramp
    if (s.spawn_speed > 1 || s.move_speed > 1) {
      if (s.ramp_timer >= 0) s.ramp_timer -= 1;
      else {
        if (s.move_speed > 1 && (s.ramp_index & 1)) s.move_speed -= 1;
        if (s.spawn_speed > 1) s.spawn_speed -= 1;
        s.ramp_index += 1;
        s.ramp_timer = RAMP_INTERVAL;
      }
    }
    s.time += 1;
    done = terminal || s.time >= max_steps;
    s.terminal = done;
  }

  PQN_HD static void obs_bits_mem(const State& s, uint32_t* o, int stride) {
    obs_set_bit(o, stride, (s.player_y * 10 + s.player_x) * OBS_C + 0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (!s.efill[e]) continue;
      const int y = e + 1;
      obs_set_bit(o, stride, (y * 10 + s.ex[e]) * OBS_C + (s.egold[e] ? 3 : 1));
      const int back = s.elr[e] ? s.ex[e] - 1 : s.ex[e] + 1;
      if (back >= 0 && back <= 9) obs_set_bit(o, stride, (y * 10 + back) * OBS_C + 2);
    }
  }
};

}  // namespace pqn
