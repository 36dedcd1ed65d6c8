"""NumPy restatement of the gymnax==0.0.6 environments the PQN hot path steps.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: gymnax is a
third-party dependency of the reference (``pyproject.toml:51``), absent from
``/root/reference`` and not installable here.  This file restates the published
algorithm of

* ``gymnax/environments/environment.py``  ``Environment.step`` (split key,
  evaluate step_env AND reset_env, select on done) / ``Environment.reset``;
* ``gymnax/environments/minatar/breakout.py``  (``MinBreakout``);
* ``gymnax/environments/classic_control/cartpole.py`` (``CartPole``),
  ``acrobot.py`` (``Acrobot``);
* ``gymnax/wrappers/purerl.py``  ``LogWrapper`` / ``FlattenObservationWrapper``
  (an in-tree copy of the LogWrapper arithmetic is
  ``purejaxql/utils/craftax_wrappers.py:161-200``).

Reference call sites: ``purejaxql/pqn_minatar.py:103-112,198-200`` and
``purejaxql/pqn_gymnax.py:92-104,192-194``.

All functions are vectorised over a leading batch of envs: ``key`` is
``uint32[N, 2]``, state is a dict of ``[N, ...]`` arrays, ``action`` ``int32[N]``.
"""
from __future__ import annotations

import numpy as np

from . import jax_prng as jr

F32 = np.float32
I32 = np.int32


# --------------------------------------------------------------------------- #
# MinAtar Breakout  (gymnax/environments/minatar/breakout.py)
# --------------------------------------------------------------------------- #
class Breakout:
    name = "Breakout-MinAtar"
    obs_shape = (10, 10, 4)
    num_actions = 3                       # minimal action set [0, 1, 3] = noop/left/right
    action_set = np.array([0, 1, 3], dtype=I32)
    max_steps_in_episode = 1000
    state_fields = ("ball_y", "ball_x", "ball_dir", "pos", "brick_map", "strike",
                    "last_y", "last_x", "time", "terminal")

    def reset_env(self, key):
        n = key.shape[0]
        # ball_start = jax.random.choice(key, jnp.array([0, 1]), shape=())
        start = jr.choice_index(key, 2, ())
        ball_x = np.where(start == 0, 0, 9).astype(I32)
        ball_dir = np.where(start == 0, 2, 3).astype(I32)
        brick = np.zeros((n, 10, 10), dtype=F32)
        brick[:, 1:4, :] = 1.0
        state = dict(
            ball_y=np.full(n, 3, I32), ball_x=ball_x, ball_dir=ball_dir,
            pos=np.full(n, 4, I32), brick_map=brick,
            strike=np.zeros(n, bool), last_y=np.full(n, 3, I32), last_x=ball_x.copy(),
            time=np.zeros(n, I32), terminal=np.zeros(n, bool),
        )
        return self.get_obs(state), state

    def get_obs(self, s):
        n = s["pos"].shape[0]
        idx = np.arange(n)
        obs = np.zeros((n, 10, 10, 4), dtype=bool)
        obs[idx, s["ball_y"], s["ball_x"], 1] = True
        obs[idx, 9, s["pos"], 0] = True
        obs[idx, s["last_y"], s["last_x"], 2] = True
        obs[:, :, :, 3] = s["brick_map"] != 0
        return obs.astype(F32)

    def step_env(self, key, s, action):
        n = action.shape[0]
        idx = np.arange(n)
        a = self.action_set[action]
        # ---- step_agent --------------------------------------------------
        pos = np.where(a == 1, np.maximum(0, s["pos"] - 1),
                       np.where(a == 3, np.minimum(9, s["pos"] + 1), s["pos"])).astype(I32)
        last_x = s["ball_x"].copy()
        last_y = s["ball_y"].copy()
        d = s["ball_dir"]
        new_x = np.where((d == 0) | (d == 3), s["ball_x"] - 1, s["ball_x"] + 1).astype(I32)
        new_y = np.where((d == 0) | (d == 1), s["ball_y"] - 1, s["ball_y"] + 1).astype(I32)
        border_x = (new_x < 0) | (new_x > 9)
        new_x = np.where(border_x, np.where(new_x < 0, 0, 9), new_x).astype(I32)
        ball_dir = np.where(border_x, np.array([1, 0, 3, 2], I32)[d], d).astype(I32)
        # ---- step_ball_brick --------------------------------------------
        border_y = new_y < 0
        new_y = np.where(border_y, 0, new_y).astype(I32)
        flip_v = np.array([3, 2, 1, 0], I32)
        ball_dir = np.where(border_y, flip_v[ball_dir], ball_dir).astype(I32)
        brick = s["brick_map"].copy()
        strike_toggle = (~border_y) & (brick[idx, new_y, new_x] == 1)
        strike_bool = (~s["strike"]) & strike_toggle
        reward = strike_bool.astype(F32)
        hit = np.nonzero(strike_bool)[0]
        brick[hit, new_y[hit], new_x[hit]] = 0.0
        new_y = np.where(strike_bool, last_y, new_y).astype(I32)
        ball_dir = np.where(strike_bool, flip_v[ball_dir], ball_dir).astype(I32)
        brick_cond = (~strike_toggle) & (new_y == 9)
        spawn = brick_cond & (np.count_nonzero(brick.reshape(n, -1), axis=1) == 0)
        brick[spawn, 1:4, :] = 1.0
        redirect1 = brick_cond & (s["ball_x"] == pos)
        ball_dir = np.where(redirect1, flip_v[ball_dir], ball_dir).astype(I32)
        new_y = np.where(redirect1, last_y, new_y).astype(I32)
        redirect2 = brick_cond & (~redirect1) & (new_x == pos)
        ball_dir = np.where(redirect2, np.array([2, 3, 0, 1], I32)[ball_dir], ball_dir).astype(I32)
        new_y = np.where(redirect2, last_y, new_y).astype(I32)
        terminal = brick_cond & (~redirect1) & (~redirect2)
        strike = strike_toggle
        time = (s["time"] + 1).astype(I32)
        done = terminal | (time >= self.max_steps_in_episode)
        ns = dict(ball_y=new_y, ball_x=new_x, ball_dir=ball_dir, pos=pos, brick_map=brick,
                  strike=strike, last_y=last_y, last_x=last_x, time=time, terminal=done)
        info = {"discount": np.where(done, F32(0.0), F32(1.0)).astype(F32)}
        return self.get_obs(ns), ns, reward, done, info


# --------------------------------------------------------------------------- #
# MinAtar Freeway  (gymnax/environments/minatar/freeway.py)
# --------------------------------------------------------------------------- #
class Freeway:
    """Restated from gymnax 0.0.6 `MinFreeway` (a JAX port of MinAtar freeway.py): chicken at column 4,
    8 cars [x, y, timer, signed speed]; cars are re-randomised (speed 1..5, direction) at reset and
    whenever the chicken reaches the top; time-limit termination only (2500 steps)."""
    name = "Freeway-MinAtar"
    obs_shape = (10, 10, 7)
    num_actions = 3                        # minimal action set [0, 2, 4] = noop/up/down
    action_set = np.array([0, 2, 4], dtype=I32)
    max_steps_in_episode = 2500
    player_speed = 3
    state_fields = ("pos", "cars", "move_timer", "time", "terminal")

    @staticmethod
    def _draw(key):
        ks = jr.split(key, 2)
        speeds = jr.randint(ks[:, 0], (8,), 1, 6)                  # jax.random.randint(key_speed, (8,), 1, 6)
        dirs = jr.choice_index(ks[:, 1], 2, (8,)) * 2 - 1           # jax.random.choice(key_dirs, [-1, 1], (8,))
        return (speeds * dirs).astype(I32)

    def reset_env(self, key):
        n = key.shape[0]
        sp = self._draw(key)
        cars = np.zeros((n, 8, 4), I32)
        cars[:, :, 1] = np.arange(1, 9)
        cars[:, :, 2] = np.abs(sp)
        cars[:, :, 3] = sp
        s = dict(pos=np.full(n, 9, I32), cars=cars, move_timer=np.full(n, self.player_speed, I32),
                 time=np.zeros(n, I32), terminal=np.zeros(n, bool))
        return self.get_obs(s), s

    def get_obs(self, s):
        n = s["pos"].shape[0]
        idx = np.arange(n)
        obs = np.zeros((n, 10, 10, 7), bool)
        obs[idx, s["pos"], 4, 0] = True
        for c in range(8):
            x, y, spd = s["cars"][:, c, 0], s["cars"][:, c, 1], s["cars"][:, c, 3]
            obs[idx, y, x, 1] = True
            back = np.where(spd > 0, x - 1, x + 1)
            back = np.where(back < 0, 9, back)
            back = np.where(back > 9, 0, back)
            obs[idx, y, back, 1 + np.abs(spd)] = True               # |speed| 1..5 -> trail channels 2..6
        return obs.astype(F32)

    def step_env(self, key, s, action):
        a = self.action_set[action]
        pos, mt = s["pos"].copy(), s["move_timer"].copy()
        up = (a == 2) & (mt == 0)
        down = (a == 4) & (mt == 0)
        mt = np.where(up | down, self.player_speed, mt).astype(I32)
        pos = np.where(up, np.maximum(0, pos - 1), np.where(down, np.minimum(9, pos + 1), pos)).astype(I32)
        win = pos == 0
        reward = win.astype(F32)
        pos = np.where(win, 9, pos).astype(I32)
        sp = self._draw(key)                                          # always sampled, applied on win
        cars = s["cars"].copy()
        cars[:, :, 2] = np.where(win[:, None], np.abs(sp), cars[:, :, 2])
        cars[:, :, 3] = np.where(win[:, None], sp, cars[:, :, 3])
        for c in range(8):                                            # sequential: pos carries across cars
            hit = (cars[:, c, 0] == 4) & (cars[:, c, 1] == pos)
            pos = np.where(hit, 9, pos).astype(I32)
            mv = cars[:, c, 2] == 0
            nx = cars[:, c, 0] + np.where(cars[:, c, 3] > 0, 1, -1)
            nx = np.where(nx < 0, 9, np.where(nx > 9, 0, nx))
            cars[:, c, 0] = np.where(mv, nx, cars[:, c, 0])
            cars[:, c, 2] = np.where(mv, np.abs(cars[:, c, 3]), cars[:, c, 2] - 1)
            hit2 = mv & (cars[:, c, 0] == 4) & (cars[:, c, 1] == pos)
            pos = np.where(hit2, 9, pos).astype(I32)
        mt = (mt - (mt > 0)).astype(I32)
        time = (s["time"] + 1).astype(I32)
        done = time >= self.max_steps_in_episode
        ns = dict(pos=pos, cars=cars.astype(I32), move_timer=mt, time=time, terminal=done)
        info = {"discount": np.where(done, F32(0.0), F32(1.0)).astype(F32)}
        return self.get_obs(ns), ns, reward, done, info


# --------------------------------------------------------------------------- #
# MinAtar SpaceInvaders  (gymnax/environments/minatar/space_invaders.py)
# --------------------------------------------------------------------------- #
class SpaceInvaders:
    """Restated from gymnax 0.0.6 `MinSpaceInvaders` (JAX port of MinAtar space_invaders.py).  Fully
    deterministic: neither step nor reset consumes randomness."""
    name = "SpaceInvaders-MinAtar"
    obs_shape = (10, 10, 6)
    num_actions = 4                        # minimal action set [0, 1, 3, 5] = noop/left/right/fire
    action_set = np.array([0, 1, 3, 5], dtype=I32)
    max_steps_in_episode = 1000
    shot_cool_down = 5
    enemy_move_interval = 12
    enemy_shot_interval = 10
    state_fields = ("pos", "f_bullet_map", "e_bullet_map", "alien_map", "alien_dir", "enemy_move_interval",
                    "alien_move_timer", "alien_shot_timer", "ramp_index", "shot_timer", "time", "terminal")

    def reset_env(self, key):
        n = key.shape[0]
        alien = np.zeros((n, 10, 10), I32)
        alien[:, 0:4, 2:8] = 1
        s = dict(pos=np.full(n, 5, I32), f_bullet_map=np.zeros((n, 10, 10), I32),
                 e_bullet_map=np.zeros((n, 10, 10), I32), alien_map=alien, alien_dir=np.full(n, -1, I32),
                 enemy_move_interval=np.full(n, self.enemy_move_interval, I32),
                 alien_move_timer=np.full(n, self.enemy_move_interval, I32),
                 alien_shot_timer=np.full(n, self.enemy_shot_interval, I32), ramp_index=np.zeros(n, I32),
                 shot_timer=np.zeros(n, I32), time=np.zeros(n, I32), terminal=np.zeros(n, bool))
        return self.get_obs(s), s

    def get_obs(self, s):
        n = s["pos"].shape[0]
        idx = np.arange(n)
        obs = np.zeros((n, 10, 10, 6), bool)
        obs[idx, 9, s["pos"], 0] = True
        al = s["alien_map"] != 0
        obs[:, :, :, 1] = al
        left = (s["alien_dir"] < 0)[:, None, None]
        obs[:, :, :, 2] = al & left
        obs[:, :, :, 3] = al & ~left
        obs[:, :, :, 4] = s["f_bullet_map"] != 0
        obs[:, :, :, 5] = s["e_bullet_map"] != 0
        return obs.astype(F32)

    def step_env(self, key, s, action):
        n = action.shape[0]
        idx = np.arange(n)
        a = self.action_set[action]
        pos = s["pos"].copy()
        fb, eb, al = s["f_bullet_map"].copy(), s["e_bullet_map"].copy(), s["alien_map"].copy()
        shot_timer = s["shot_timer"].copy()
        # ---- player action
        fire = (a == 5) & (shot_timer == 0)
        fb[idx[fire], 9, pos[fire]] = 1
        shot_timer = np.where(fire, self.shot_cool_down, shot_timer).astype(I32)
        pos = np.where(~fire & (a == 1), np.maximum(0, pos - 1), pos)
        pos = np.where(~fire & (a == 3), np.minimum(9, pos + 1), pos).astype(I32)
        # ---- bullets
        fb = np.roll(fb, -1, axis=1); fb[:, 9, :] = 0
        eb = np.roll(eb, 1, axis=1); eb[:, 0, :] = 0
        terminal = eb[idx, 9, pos] != 0
        # ---- aliens
        terminal |= al[idx, 9, pos] != 0
        move = s["alien_move_timer"] == 0
        amt = np.where(move, np.minimum(np.count_nonzero(al.reshape(n, -1), axis=1), s["enemy_move_interval"]),
                       s["alien_move_timer"]).astype(I32)
        adir = s["alien_dir"].copy()
        edge = ((al[:, :, 0].sum(1) > 0) & (adir < 0)) | ((al[:, :, 9].sum(1) > 0) & (adir > 0))
        rev = move & edge
        terminal |= rev & (al[:, 9, :].sum(1) > 0)
        adir = np.where(rev, -adir, adir).astype(I32)
        al_down = np.roll(al, 1, axis=1)
        al_left = np.roll(al, -1, axis=2)
        al_right = np.roll(al, 1, axis=2)
        side = move & ~edge
        al = np.where(rev[:, None, None], al_down, al)
        al = np.where((side & (s["alien_dir"] < 0))[:, None, None], al_left, al)
        al = np.where((side & (s["alien_dir"] > 0))[:, None, None], al_right, al)
        terminal |= move & (al[idx, 9, pos] != 0)
        # ---- alien shot from the alien nearest to the cannon (ties: lower column first)
        shoot = s["alien_shot_timer"] == 0
        ast = np.where(shoot, self.enemy_shot_interval, s["alien_shot_timer"]).astype(I32)
        cols = np.arange(10)
        for i in np.nonzero(shoot)[0]:
            order = np.argsort(np.abs(cols - pos[i]), kind="stable")
            for c in order:
                if al[i, :, c].sum() > 0:
                    eb[i, np.max(np.nonzero(al[i, :, c])[0]), c] = 1
                    break
        # ---- kills
        kill = (al != 0) & (fb != 0)
        reward = kill.reshape(n, -1).sum(1).astype(F32)
        al = np.where(kill, 0, al); fb = np.where(kill, 0, fb)
        # ---- timers, ramping, respawn
        shot_timer = (shot_timer - (shot_timer > 0)).astype(I32)
        amt = (amt - 1).astype(I32)
        ast = (ast - 1).astype(I32)
        empty = np.count_nonzero(al.reshape(n, -1), axis=1) == 0
        ramp = empty & (s["enemy_move_interval"] > 6)
        emi = np.where(ramp, s["enemy_move_interval"] - 1, s["enemy_move_interval"]).astype(I32)
        ridx = np.where(ramp, s["ramp_index"] + 1, s["ramp_index"]).astype(I32)
        al[empty, 0:4, 2:8] = 1
        time = (s["time"] + 1).astype(I32)
        done = terminal | (time >= self.max_steps_in_episode)
        ns = dict(pos=pos, f_bullet_map=fb.astype(I32), e_bullet_map=eb.astype(I32), alien_map=al.astype(I32),
                  alien_dir=adir, enemy_move_interval=emi, alien_move_timer=amt, alien_shot_timer=ast,
                  ramp_index=ridx, shot_timer=shot_timer, time=time, terminal=done)
        info = {"discount": np.where(done, F32(0.0), F32(1.0)).astype(F32)}
        return self.get_obs(ns), ns, reward, done, info


# --------------------------------------------------------------------------- #
# MinAtar Asterix  (gymnax/environments/minatar/asterix.py)
# --------------------------------------------------------------------------- #
def _choice_p(key, p):
    """``jax.random.choice(key, len(p), p=p)`` per row: cumsum, r = total*(1-uniform), searchsorted-left (f32)."""
    cum = np.cumsum(p.astype(F32), axis=-1, dtype=F32)
    r = (cum[:, -1] * (F32(1.0) - jr.uniform(key, ()))).astype(F32)
    with np.errstate(invalid="ignore"):
        return (cum < r[:, None]).sum(-1).astype(I32)


class Asterix:
    """Restated from MinAtar asterix.py in gymnax 0.0.6's conventions (entities [x, y=slot+1, lr, is_gold,
    filled]; every step splits its key three ways for direction / treasure / slot draws).  The exact RNG
    recipe of gymnax's `spawn_entity` could not be checked against a live install (parity unpinned)."""
    name = "Asterix-MinAtar"
    obs_shape = (10, 10, 4)
    num_actions = 5                        # minimal action set [0..4] = noop/left/up/right/down
    max_steps_in_episode = 1000
    ramp_interval, init_spawn_speed, init_move_interval = 100, 10, 5
    state_fields = ("player_x", "player_y", "shot_timer", "spawn_speed", "spawn_timer", "move_speed", "move_timer",
                    "ramp_timer", "ramp_index", "entities", "time", "terminal")

    def reset_env(self, key):
        n = key.shape[0]
        z = lambda v: np.full(n, v, I32)
        s = dict(player_x=z(5), player_y=z(5), shot_timer=z(0), spawn_speed=z(self.init_spawn_speed),
                 spawn_timer=z(self.init_spawn_speed), move_speed=z(self.init_move_interval),
                 move_timer=z(self.init_move_interval), ramp_timer=z(self.ramp_interval), ramp_index=z(0),
                 entities=np.zeros((n, 8, 5), I32), time=z(0), terminal=np.zeros(n, bool))
        return self.get_obs(s), s

    def get_obs(self, s):
        n = s["player_x"].shape[0]
        idx = np.arange(n)
        obs = np.zeros((n, 10, 10, 4), bool)
        obs[idx, s["player_y"], s["player_x"], 0] = True
        for e in range(8):
            ent = s["entities"][:, e]
            on = ent[:, 4] == 1
            ch = np.where(ent[:, 3] == 1, 3, 1)
            obs[idx[on], ent[on, 1], ent[on, 0], ch[on]] = True
            back = np.where(ent[:, 2] == 1, ent[:, 0] - 1, ent[:, 0] + 1)
            ok = on & (back >= 0) & (back <= 9)
            obs[idx[ok], ent[ok, 1], back[ok], 2] = True
        return obs.astype(F32)

    def step_env(self, key, s, action):
        n = action.shape[0]
        ent = s["entities"].copy()
        ks = jr.split(key, 3)
        lr = (1 - jr.randint(ks[:, 0], (), 0, 2)).astype(I32)                 # choice(key_lr, [1, 0])
        is_gold = (1 - _choice_p(ks[:, 1], np.tile(np.array([1 / 3, 2 / 3], F32), (n, 1)))).astype(I32)
        free = (ent[:, :, 4] == 0)
        nfree = free.sum(1)
        with np.errstate(invalid="ignore", divide="ignore"):
            pslot = free.astype(F32) / nfree.astype(F32)[:, None]
        slot = np.minimum(_choice_p(ks[:, 2], pslot), 7)
        do_spawn = (s["spawn_timer"] == 0) & (nfree > 0)
        for i in np.nonzero(do_spawn)[0]:
            ent[i, slot[i]] = (0 if lr[i] else 9, slot[i] + 1, lr[i], is_gold[i], 1)
        spawn_timer = np.where(s["spawn_timer"] == 0, s["spawn_speed"], s["spawn_timer"]).astype(I32)
        # ---- player
        px, py = s["player_x"].copy(), s["player_y"].copy()
        px = np.where(action == 1, np.maximum(0, px - 1), np.where(action == 3, np.minimum(9, px + 1), px)).astype(I32)
        py = np.where(action == 2, np.maximum(1, py - 1), np.where(action == 4, np.minimum(8, py + 1), py)).astype(I32)
        reward = np.zeros(n, F32)
        terminal = np.zeros(n, bool)

        def collide():
            nonlocal reward, terminal
            for e in range(8):
                hit = (ent[:, e, 4] == 1) & (ent[:, e, 0] == px) & (ent[:, e, 1] == py)
                gold = hit & (ent[:, e, 3] == 1)
                reward = reward + gold.astype(F32)
                terminal = terminal | (hit & ~gold)
                ent[gold, e] = 0
        collide()
        move = s["move_timer"] == 0
        move_timer = np.where(move, s["move_speed"], s["move_timer"]).astype(I32)
        for e in range(8):
            on = move & (ent[:, e, 4] == 1)
            nx = ent[:, e, 0] + np.where(ent[:, e, 2] == 1, 1, -1)
            ent[:, e, 0] = np.where(on, nx, ent[:, e, 0])
            out = on & ((nx < 0) | (nx > 9))
            ent[out, e] = 0
        mv_saved = move
        if mv_saved.any():
            r0, t0 = reward.copy(), terminal.copy()
            ent_before = ent.copy()
            collide()
            # the second collision pass only applies where the entities moved
            reward = np.where(mv_saved, reward, r0)
            terminal = np.where(mv_saved, terminal, t0)
            ent = np.where(mv_saved[:, None, None], ent, ent_before)
        spawn_timer = (spawn_timer - 1).astype(I32)
        move_timer = (move_timer - 1).astype(I32)
        # ---- difficulty ramp
        ss, ms, rt, ri = s["spawn_speed"].copy(), s["move_speed"].copy(), s["ramp_timer"].copy(), s["ramp_index"].copy()
        active = (ss > 1) | (ms > 1)
        tick = active & (rt >= 0)
        fire = active & ~tick
        ms = np.where(fire & (ms > 1) & (ri % 2 == 1), ms - 1, ms).astype(I32)
        ss = np.where(fire & (ss > 1), ss - 1, ss).astype(I32)
        ri = np.where(fire, ri + 1, ri).astype(I32)
        rt = np.where(tick, rt - 1, np.where(fire, self.ramp_interval, rt)).astype(I32)
        time = (s["time"] + 1).astype(I32)
        done = terminal | (time >= self.max_steps_in_episode)
        ns = dict(player_x=px, player_y=py, shot_timer=s["shot_timer"].copy(), spawn_speed=ss, spawn_timer=spawn_timer,
                  move_speed=ms, move_timer=move_timer, ramp_timer=rt, ramp_index=ri, entities=ent.astype(I32),
                  time=time, terminal=done)
        info = {"discount": np.where(done, F32(0.0), F32(1.0)).astype(F32)}
        return self.get_obs(ns), ns, reward, done, info


# --------------------------------------------------------------------------- #
# CartPole-v1  (gymnax/environments/classic_control/cartpole.py)
# --------------------------------------------------------------------------- #
class CartPole:
    name = "CartPole-v1"
    obs_shape = (4,)
    num_actions = 2
    max_steps_in_episode = 500
    state_fields = ("x", "x_dot", "theta", "theta_dot", "time")
    gravity = F32(9.8)
    masscart = F32(1.0)
    masspole = F32(0.1)
    total_mass = F32(1.0 + 0.1)
    length = F32(0.5)
    polemass_length = F32(0.05)
    force_mag = F32(10.0)
    tau = F32(0.02)
    theta_threshold_radians = F32(12 * 2 * np.pi / 360)
    x_threshold = F32(2.4)

    def is_terminal(self, s):
        d1 = (s["x"] < -self.x_threshold) | (s["x"] > self.x_threshold)
        d2 = (s["theta"] < -self.theta_threshold_radians) | (s["theta"] > self.theta_threshold_radians)
        return d1 | d2 | (s["time"] >= self.max_steps_in_episode)

    def get_obs(self, s):
        return np.stack([s["x"], s["x_dot"], s["theta"], s["theta_dot"]], axis=-1).astype(F32)

    def reset_env(self, key):
        u = jr.uniform(key, (4,), -0.05, 0.05)
        s = dict(x=u[:, 0], x_dot=u[:, 1], theta=u[:, 2], theta_dot=u[:, 3],
                 time=np.zeros(key.shape[0], I32))
        return self.get_obs(s), s

    def step_env(self, key, s, action):
        prev_terminal = self.is_terminal(s)
        af = action.astype(F32)
        force = self.force_mag * af - self.force_mag * (F32(1.0) - af)
        costheta = np.cos(s["theta"]).astype(F32)
        sintheta = np.sin(s["theta"]).astype(F32)
        temp = (force + self.polemass_length * s["theta_dot"] ** 2 * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (F32(4.0 / 3.0) - self.masspole * costheta ** 2 / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        x = s["x"] + self.tau * s["x_dot"]
        x_dot = s["x_dot"] + self.tau * xacc
        theta = s["theta"] + self.tau * s["theta_dot"]
        theta_dot = s["theta_dot"] + self.tau * thetaacc
        reward = (F32(1.0) - prev_terminal.astype(F32)).astype(F32)
        ns = dict(x=x.astype(F32), x_dot=x_dot.astype(F32), theta=theta.astype(F32),
                  theta_dot=theta_dot.astype(F32), time=(s["time"] + 1).astype(I32))
        done = self.is_terminal(ns)
        info = {"discount": np.where(done, F32(0.0), F32(1.0)).astype(F32)}
        return self.get_obs(ns), ns, reward, done, info


# --------------------------------------------------------------------------- #
# Acrobot-v1  (gymnax/environments/classic_control/acrobot.py)
# --------------------------------------------------------------------------- #
class Acrobot:
    name = "Acrobot-v1"
    obs_shape = (6,)
    num_actions = 3
    max_steps_in_episode = 500
    state_fields = ("joint_angle1", "joint_angle2", "velocity_1", "velocity_2", "time")
    dt = F32(0.2)
    max_vel_1 = F32(4 * np.pi)
    max_vel_2 = F32(9 * np.pi)
    torques = np.array([-1.0, 0.0, 1.0], F32)
    torque_noise_max = F32(0.0)

    @staticmethod
    def _dsdt(s):
        m1 = m2 = l1 = F32(1.0)
        lc1 = lc2 = F32(0.5)
        I1 = I2 = F32(1.0)
        g = F32(9.8)
        pi = F32(np.pi)
        th1, th2, d1_, d2_, a = s
        d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + F32(2) * l1 * lc2 * np.cos(th2)) + I1 + I2
        d2 = m2 * (lc2 ** 2 + l1 * lc2 * np.cos(th2)) + I2
        phi2 = m2 * lc2 * g * np.cos(th1 + th2 - pi / F32(2.0))
        phi1 = (-m2 * l1 * lc2 * d2_ ** 2 * np.sin(th2)
                - F32(2) * m2 * l1 * lc2 * d2_ * d1_ * np.sin(th2)
                + (m1 * lc1 + m2 * l1) * g * np.cos(th1 - pi / F32(2))
                + phi2)
        dd2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * d1_ ** 2 * np.sin(th2) - phi2) / (
            m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
        dd1 = -(a + d2 * dd2 + phi1) / d1
        z = np.zeros_like(th1)
        return [x.astype(F32) for x in (d1_, d2_, dd1, dd2, z)]

    def _rk4(self, y0):
        dt2 = self.dt / F32(2.0)
        add = lambda y, h, k: [(yi + h * ki).astype(F32) for yi, ki in zip(y, k)]
        k1 = self._dsdt(y0)
        k2 = self._dsdt(add(y0, dt2, k1))
        k3 = self._dsdt(add(y0, dt2, k2))
        k4 = self._dsdt(add(y0, self.dt, k3))
        return [(y + self.dt / F32(6.0) * (a + F32(2) * b + F32(2) * c + d)).astype(F32)
                for y, a, b, c, d in zip(y0, k1, k2, k3, k4)]

    @staticmethod
    def _wrap(x, m, M):
        diff = F32(M - m)
        go_up = x < m
        go_down = x >= M
        how_often = (go_up * np.ceil((m - x) / diff) + go_down * np.floor((x - m) / diff)).astype(F32)
        return (x - how_often * diff * go_down + how_often * diff * go_up).astype(F32)

    def get_obs(self, s):
        return np.stack([np.cos(s["joint_angle1"]), np.sin(s["joint_angle1"]),
                         np.cos(s["joint_angle2"]), np.sin(s["joint_angle2"]),
                         s["velocity_1"], s["velocity_2"]], axis=-1).astype(F32)

    def reset_env(self, key):
        u = jr.uniform(key, (4,), -0.1, 0.1)
        s = dict(joint_angle1=u[:, 0], joint_angle2=u[:, 1], velocity_1=u[:, 2],
                 velocity_2=u[:, 3], time=np.zeros(key.shape[0], I32))
        return self.get_obs(s), s

    def step_env(self, key, s, action):
        torque = self.torques[action]
        torque = (torque + jr.uniform(key, (), -self.torque_noise_max, self.torque_noise_max)).astype(F32)
        ns = self._rk4([s["joint_angle1"], s["joint_angle2"], s["velocity_1"], s["velocity_2"], torque])
        pi = F32(np.pi)
        a1 = self._wrap(ns[0], -pi, pi)
        a2 = self._wrap(ns[1], -pi, pi)
        v1 = np.clip(ns[2], -self.max_vel_1, self.max_vel_1).astype(F32)
        v2 = np.clip(ns[3], -self.max_vel_2, self.max_vel_2).astype(F32)
        time = (s["time"] + 1).astype(I32)
        done_angle = (-np.cos(a1) - np.cos(a2 + a1)) > F32(1.0)
        done = done_angle | (time >= self.max_steps_in_episode)
        reward = (F32(-1.0) * (F32(1.0) - done_angle.astype(F32))).astype(F32)
        st = dict(joint_angle1=a1, joint_angle2=a2, velocity_1=v1, velocity_2=v2, time=time)
        info = {"discount": np.where(done, F32(0.0), F32(1.0)).astype(F32)}
        return self.get_obs(st), st, reward, done, info


# --------------------------------------------------------------------------- #
# gymnax core: Environment.step auto-reset, wrappers
# --------------------------------------------------------------------------- #
def _select(done, a_re, a_st):
    d = done.reshape(done.shape + (1,) * (a_st.ndim - 1))
    return np.where(d, a_re, a_st).astype(a_st.dtype)


class Environment:
    """``gymnax.environments.environment.Environment`` protocol, batched."""

    def __init__(self, core, flatten=False):
        self.core = core
        self.flatten = flatten            # FlattenObservationWrapper (pqn_gymnax.py:93)

    @property
    def num_actions(self):
        return self.core.num_actions

    @property
    def obs_shape(self):
        s = self.core.obs_shape
        return (int(np.prod(s)),) if self.flatten else s

    def _f(self, obs):
        return obs.reshape(obs.shape[0], -1) if self.flatten else obs

    def reset(self, key):
        obs, st = self.core.reset_env(np.asarray(key, np.uint32))
        return self._f(obs), st

    def step(self, key, state, action):
        ks = jr.split(np.asarray(key, np.uint32), 2)
        k_step, k_reset = ks[:, 0], ks[:, 1]
        obs_st, state_st, reward, done, info = self.core.step_env(k_step, state, np.asarray(action, I32))
        obs_re, state_re = self.core.reset_env(k_reset)
        state = {k: _select(done, state_re[k], state_st[k]) for k in state_st}
        obs = _select(done, obs_re, obs_st)
        return self._f(obs), state, reward.astype(F32), done, info


class LogWrapper:
    """``gymnax.wrappers.purerl.LogWrapper`` (arithmetic as in
    ``purejaxql/utils/craftax_wrappers.py:173-200``).  The log fields travel in
    the same state dict under ``log_*`` keys."""

    LOG = ("log_episode_returns", "log_episode_lengths", "log_returned_episode_returns",
           "log_returned_episode_lengths", "log_timestep")

    def __init__(self, env: Environment):
        self.env = env
        self.num_actions = env.num_actions
        self.obs_shape = env.obs_shape

    def reset(self, key):
        obs, st = self.env.reset(key)
        n = obs.shape[0]
        st = dict(st)
        st["log_episode_returns"] = np.zeros(n, F32)
        st["log_episode_lengths"] = np.zeros(n, I32)
        st["log_returned_episode_returns"] = np.zeros(n, F32)
        st["log_returned_episode_lengths"] = np.zeros(n, I32)
        st["log_timestep"] = np.zeros(n, I32)
        return obs, st

    def step(self, key, state, action):
        inner = {k: v for k, v in state.items() if not k.startswith("log_")}
        obs, env_state, reward, done, info = self.env.step(key, inner, action)
        d_f = done.astype(F32)
        d_i = done.astype(I32)
        new_ret = (state["log_episode_returns"] + reward).astype(F32)
        new_len = (state["log_episode_lengths"] + 1).astype(I32)
        st = dict(env_state)
        st["log_episode_returns"] = (new_ret * (F32(1) - d_f)).astype(F32)
        st["log_episode_lengths"] = (new_len * (1 - d_i)).astype(I32)
        st["log_returned_episode_returns"] = (
            state["log_returned_episode_returns"] * (F32(1) - d_f) + new_ret * d_f).astype(F32)
        st["log_returned_episode_lengths"] = (
            state["log_returned_episode_lengths"] * (1 - d_i) + new_len * d_i).astype(I32)
        st["log_timestep"] = (state["log_timestep"] + 1).astype(I32)
        info = dict(info)
        info["returned_episode_returns"] = st["log_returned_episode_returns"]
        info["returned_episode_lengths"] = st["log_returned_episode_lengths"]
        info["timestep"] = st["log_timestep"]
        info["returned_episode"] = done
        return obs, st, reward, done, info


_REGISTRY = {}


def register(cls):
    _REGISTRY[cls.name] = cls
    return cls


for _c in (Breakout, Asterix, Freeway, SpaceInvaders, CartPole, Acrobot):
    register(_c)


def make(env_name: str, flatten: bool = False, log: bool = True):
    """``gymnax.make(env_name)`` + wrapper composition used by the reference
    (``LogWrapper(env)`` in pqn_minatar.py:103-104;
    ``LogWrapper(FlattenObservationWrapper(env))`` in pqn_gymnax.py:92-94)."""
    core = _REGISTRY[env_name]()
    env = Environment(core, flatten=flatten)
    return LogWrapper(env) if log else env
