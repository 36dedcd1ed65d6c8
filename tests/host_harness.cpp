// CPU logic harness — TEST INFRASTRUCTURE ONLY (built by tests/_harness.py).
// Compiles the SAME inline per-env functions the CUDA kernels use
// (purejaxql_b200/csrc/*.cuh are __host__ __device__) with g++, so that the
// env / PRNG / eps-greedy / Q(lambda) logic can be checked against oracle/ in the
// GPU-less build container before GPU minutes are spent.  The product never
// calls this; libpqn_b200.so has no host compute path.
#include <stdint.h>
#include <string.h>

#include "../purejaxql_b200/csrc/env_breakout.cuh"
#include "../purejaxql_b200/csrc/env_classic.cuh"
#include "../purejaxql_b200/csrc/env_minatar_more.cuh"
#include "../purejaxql_b200/csrc/rollout_logic.cuh"

using namespace pqn;

template <class Env>
static void obs_out(const typename Env::State& s, float* obs, int64_t i) {
  if constexpr (Env::BINARY_OBS) {
    uint32_t bits[Env::OBS_WORDS_PAD];
    if constexpr (Env::OBS_IN_REGS) {
      Env::obs_bits(s, bits);
    } else {
      for (int w = 0; w < Env::OBS_WORDS_PAD; ++w) bits[w] = 0u;
      Env::obs_bits_mem(s, bits, 1);
    }
    for (int f = 0; f < Env::OBS_DIM; ++f) obs[i * Env::OBS_DIM + f] = (float)((bits[f >> 5] >> (f & 31)) & 1u);
  } else {
    float o[Env::OBS_DIM];
    Env::obs_float(s, o);
    for (int f = 0; f < Env::OBS_DIM; ++f) obs[i * Env::OBS_DIM + f] = o[f];
  }
}

template <class Env>
static void reset_t(const uint32_t* keys, uint32_t* state, float* obs, int64_t N, int max_steps, int part) {
  for (int64_t i = 0; i < N; ++i) {
    typename Env::State s;
    Env::reset_env(Key{keys[2 * i], keys[2 * i + 1]}, part, max_steps, s);
    Env::store(s, state, N, i);
    LogState lg;
    log_reset(lg);
    log_store(lg, state, N, i, Env::CORE_WORDS);
    obs_out<Env>(s, obs, i);
  }
}

template <class Env>
static void step_t(const uint32_t* keys, uint32_t* state, const int32_t* action, float* obs, float* reward,
                   uint8_t* done, int64_t N, int max_steps, int part) {
  for (int64_t i = 0; i < N; ++i) {
    typename Env::State s;
    Env::load(s, state, N, i);
    LogState lg;
    log_load(lg, state, N, i, Env::CORE_WORDS);
    float r;
    bool d;
    env_step_full<Env>(Key{keys[2 * i], keys[2 * i + 1]}, part, max_steps, s, lg, action[i], r, d);
    Env::store(s, state, N, i);
    log_store(lg, state, N, i, Env::CORE_WORDS);
    reward[i] = r;
    done[i] = d;
    obs_out<Env>(s, obs, i);
  }
}

extern "C" {
int h_state_words(int env_id) {
  switch (env_id) {
    case ENV_BREAKOUT: return BreakoutEnv::STATE_WORDS;
    case ENV_FREEWAY: return FreewayEnv::STATE_WORDS;
    case ENV_ASTERIX: return AsterixEnv::STATE_WORDS;
    case ENV_SPACE_INVADERS: return SpaceInvadersEnv::STATE_WORDS;
    case ENV_CARTPOLE: return CartPoleEnv::STATE_WORDS;
    case ENV_ACROBOT: return AcrobotEnv::STATE_WORDS;
  }
  return -1;
}
int h_env_reset(int env_id, const uint32_t* keys, uint32_t* state, float* obs, int64_t N, int max_steps, int part) {
  switch (env_id) {
    case ENV_BREAKOUT: reset_t<BreakoutEnv>(keys, state, obs, N, max_steps, part); return 0;
    case ENV_FREEWAY: reset_t<FreewayEnv>(keys, state, obs, N, max_steps, part); return 0;
    case ENV_ASTERIX: reset_t<AsterixEnv>(keys, state, obs, N, max_steps, part); return 0;
    case ENV_SPACE_INVADERS: reset_t<SpaceInvadersEnv>(keys, state, obs, N, max_steps, part); return 0;
    case ENV_CARTPOLE: reset_t<CartPoleEnv>(keys, state, obs, N, max_steps, part); return 0;
    case ENV_ACROBOT: reset_t<AcrobotEnv>(keys, state, obs, N, max_steps, part); return 0;
  }
  return -1;
}
int h_env_step(int env_id, const uint32_t* keys, uint32_t* state, const int32_t* action, float* obs, float* reward,
               uint8_t* done, int64_t N, int max_steps, int part) {
  switch (env_id) {
    case ENV_BREAKOUT: step_t<BreakoutEnv>(keys, state, action, obs, reward, done, N, max_steps, part); return 0;
    case ENV_FREEWAY: step_t<FreewayEnv>(keys, state, action, obs, reward, done, N, max_steps, part); return 0;
    case ENV_ASTERIX: step_t<AsterixEnv>(keys, state, action, obs, reward, done, N, max_steps, part); return 0;
    case ENV_SPACE_INVADERS: step_t<SpaceInvadersEnv>(keys, state, action, obs, reward, done, N, max_steps, part); return 0;
    case ENV_CARTPOLE: step_t<CartPoleEnv>(keys, state, action, obs, reward, done, N, max_steps, part); return 0;
    case ENV_ACROBOT: step_t<AcrobotEnv>(keys, state, action, obs, reward, done, N, max_steps, part); return 0;
  }
  return -1;
}
void h_split(const uint32_t* keys, int64_t n, int num, uint32_t* out, int part) {
  for (int64_t k = 0; k < n; ++k)
    for (int j = 0; j < num; ++j) {
      Key c = split_at(Key{keys[2 * k], keys[2 * k + 1]}, num, j, part);
      out[2 * (k * num + j)] = c.k0;
      out[2 * (k * num + j) + 1] = c.k1;
    }
}
void h_split3(const uint32_t* key, uint32_t* out, int part) {
  Key a, b, c;
  split3(Key{key[0], key[1]}, part, a, b, c);
  out[0] = a.k0; out[1] = a.k1; out[2] = b.k0; out[3] = b.k1; out[4] = c.k0; out[5] = c.k1;
}
void h_bits(const uint32_t* key, int64_t len, uint32_t* out, int part) {
  for (int64_t j = 0; j < len; ++j) out[j] = bits_at(Key{key[0], key[1]}, (uint32_t)len, (uint32_t)j, part);
}
void h_eps_greedy(const uint32_t* keys, const float* q, float eps, int32_t* action, float* maxq, int64_t N, int A,
                  int part) {
  for (int64_t i = 0; i < N; ++i)
    action[i] = eps_greedy_one(Key{keys[2 * i], keys[2 * i + 1]}, q + i * A, A, eps, part, maxq[i]);
}
void h_qlambda(const float* reward, const uint8_t* done, const float* maxq, const float* q_last, float* targets, int T,
               int64_t N, int A, float gamma, float lambda) {
  for (int64_t i = 0; i < N; ++i) qlambda_one(reward, done, maxq, q_last + i * A, targets, T, N, A, gamma, lambda, i);
}
}
