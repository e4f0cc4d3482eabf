/* CPU checker for the env side of the PPO hot path -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the same gym==0.26.2 algorithm as
 * oracle/gym_restated.py (see that file's header for the upstream files and
 * the reference call sites ppo.py:66-68,87-97,110,188), vectorised the way
 * gym.vector.SyncVectorEnv is: a serial loop over envs on one core.
 * PARITY UNPINNED at the gym boundary (third-party, un-vendored, no reference
 * tests); pinned against gym_restated.py (bit-for-bit in libm mode) and
 * against NumPy's own PCG64 in tests/test_oracle_envs.py.
 *
 * trig mode 0 = host libm sin/cos (what gym calls); trig mode 1 = the
 * deterministic routine in aur_ppo_b200/csrc/det_sincos.h (the single header
 * the GPU kernels also compile), so that GPU-vs-checker transitions can be
 * compared bit-for-bit; trig mode 2 = the checker's OWN reference, independent
 * of the product: libquadmath's 113-bit sinq / cosq rounded once to double,
 * i.e. the correctly rounded value.  tests/test_oracle_envs.py replays the
 * GPU parity trajectories in modes 1 and 2 and requires identical bits, so the
 * bit-exact GPU == checker claim does not rest on the shared header alone.
 * Build with -ffp-contract=off.
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../aur_ppo_b200/csrc/det_sincos.h"

typedef unsigned __int128 u128;

/* ---- PCG64 (numpy/random/src/pcg64: setseq_128, XSL-RR 128/64) ---- */
typedef struct { u128 state, inc; } pcg64_t;
static const u128 PCG_MULT = ((u128)0x2360ED051FC65DA4ULL << 64) | 0x4385DF649FCCF645ULL;

static inline uint64_t pcg64_next(pcg64_t* g) {
  g->state = g->state * PCG_MULT + g->inc;
  uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state;
  uint64_t x = hi ^ lo;
  unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((-rot) & 63));
}
static inline double pcg64_double(pcg64_t* g) { return (double)(pcg64_next(g) >> 11) * (1.0 / 9007199254740992.0); }
/* Generator.uniform(low, high): low + (high - low) * next_double */
static inline double pcg64_uniform(pcg64_t* g, double low, double range) { return low + range * pcg64_double(g); }

static inline int obs_dim_of(int kind) { return kind == 0 ? 4 : (kind == 1 ? 3 : (kind == 3 ? 6 : 2)); }
static inline int time_limit_of(int kind) { return (kind == 0 || kind == 3) ? 500 : (kind == 4 ? 999 : 200); }

static inline void trig(int mode, double x, double* s, double* c) {
  if (mode == 0) { *s = sin(x); *c = cos(x); }
  else if (mode == 2) { *s = (double)sinq((__float128)x); *c = (double)cosq((__float128)x); }
  else aur_sincos(x, s, c);
}

void orc_sincos(int64_t n, const double* x, double* s, double* c) {
  for (int64_t i = 0; i < n; ++i) aur_sincos(x[i], &s[i], &c[i]);
}
void orc_cr_sincos(int64_t n, const double* x, double* s, double* c) {
  for (int64_t i = 0; i < n; ++i) { s[i] = (double)sinq((__float128)x[i]); c[i] = (double)cosq((__float128)x[i]); }
}
void orc_libm_sincos(int64_t n, const double* x, double* s, double* c) {
  for (int64_t i = 0; i < n; ++i) { s[i] = sin(x[i]); c[i] = cos(x[i]); }
}
void orc_pcg64_doubles(uint64_t st_hi, uint64_t st_lo, uint64_t inc_hi, uint64_t inc_lo, int64_t n, double* out) {
  pcg64_t g; g.state = ((u128)st_hi << 64) | st_lo; g.inc = ((u128)inc_hi << 64) | inc_lo;
  for (int64_t i = 0; i < n; ++i) out[i] = pcg64_double(&g);
}

/* ---- RunningMeanStd.update with batch_count == 1 (gym/wrappers/normalize.py) ---- */
static inline void rms_update1(double* mean, double* var, double* count, double x) {
  double delta = x - *mean;
  double tot = *count + 1;
  double new_mean = *mean + delta * 1 / tot;
  double m_a = *var * *count;
  double m_b = 0.0 * 1;
  double M2 = m_a + m_b + delta * delta * *count * 1 / tot;
  *mean = new_mean; *var = M2 / tot; *count = tot;
}

/* ---- Acrobot-v1 (gym/envs/classic_control/acrobot.py: AcrobotEnv._dsdt with book_or_nips = "book", rk4, wrap, bound) ---- */
static void acrobot_dsdt(int mode, const double* y, double a, double* k) {
  const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 0.5, I1 = 1.0, I2 = 1.0, g = 9.8;
  const double pi = 3.141592653589793;
  double theta1 = y[0], theta2 = y[1], dtheta1 = y[2], dtheta2 = y[3];
  double sin2, cos2, sd, c12, c1;
  trig(mode, theta2, &sin2, &cos2);
  trig(mode, theta1 + theta2 - pi / 2.0, &sd, &c12);
  trig(mode, theta1 - pi / 2, &sd, &c1);
  double d1 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2 + 2 * l1 * lc2 * cos2) + I1 + I2;
  double d2 = m2 * (lc2 * lc2 + l1 * lc2 * cos2) + I2;
  double phi2 = m2 * lc2 * g * c12;
  double phi1 = -m2 * l1 * lc2 * (dtheta2 * dtheta2) * sin2 - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * sin2 +
                (m1 * lc1 + m2 * l1) * g * c1 + phi2;
  double ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * (dtheta1 * dtheta1) * sin2 - phi2) /
                    (m2 * (lc2 * lc2) + I2 - (d2 * d2) / d1);
  double ddtheta1 = -(d2 * ddtheta2 + phi1) / d1;
  k[0] = dtheta1; k[1] = dtheta2; k[2] = ddtheta1; k[3] = ddtheta2;
}
/* one env.step of Acrobot: st[4] updated in place, returns terminated */
static int acrobot_step(int mode, double* st, int action) {
  const double pi = 3.141592653589793, dt = 0.2 - 0, dt2 = dt / 2.0;
  double torque = (double)(action - 1);     /* AVAIL_TORQUE = [-1., 0., +1] */
  double k1[4], k2[4], k3[4], k4[4], y[4], ns[4];
  acrobot_dsdt(mode, st, torque, k1);
  for (int i = 0; i < 4; ++i) y[i] = st[i] + dt2 * k1[i];
  acrobot_dsdt(mode, y, torque, k2);
  for (int i = 0; i < 4; ++i) y[i] = st[i] + dt2 * k2[i];
  acrobot_dsdt(mode, y, torque, k3);
  for (int i = 0; i < 4; ++i) y[i] = st[i] + dt * k3[i];
  acrobot_dsdt(mode, y, torque, k4);
  for (int i = 0; i < 4; ++i) ns[i] = st[i] + dt / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
  double diff = pi - (-pi);
  for (int i = 0; i < 2; ++i) {
    while (ns[i] > pi) ns[i] = ns[i] - diff;
    while (ns[i] < -pi) ns[i] = ns[i] + diff;
  }
  double mv1 = 4 * pi, mv2 = 9 * pi;
  ns[2] = ns[2] < -mv1 ? -mv1 : ns[2]; ns[2] = ns[2] > mv1 ? mv1 : ns[2];
  ns[3] = ns[3] < -mv2 ? -mv2 : ns[3]; ns[3] = ns[3] > mv2 ? mv2 : ns[3];
  for (int i = 0; i < 4; ++i) st[i] = ns[i];
  double sa, ca, sb, cb;
  trig(mode, st[0], &sa, &ca);
  trig(mode, st[1] + st[0], &sb, &cb);
  return (-ca - cb) > 1.0;
}
static void acrobot_raw_obs(int mode, const double* st, double* raw) {
  double sa, ca, sb, cb;
  trig(mode, st[0], &sa, &ca);
  trig(mode, st[1], &sb, &cb);
  raw[0] = (double)(float)ca; raw[1] = (double)(float)sa; raw[2] = (double)(float)cb; raw[3] = (double)(float)sb;
  raw[4] = (double)(float)st[2]; raw[5] = (double)(float)st[3];
}

/* ---- vector env ---- */
typedef struct {
  int kind;        /* 0 CartPole-v1, 1 Pendulum-v1, 2 MountainCar-v0, 3 Acrobot-v1, 4 MountainCarContinuous-v0 */
  int wrappers;    /* 1 = the reference's continuous wrapper stack (ppo.py:92-97) */
  int trig_mode;
  int64_t n;
  double gamma;    /* NormalizeReward gamma (0.99) */
  pcg64_t* rng;
  double* phys;    /* [n][4] or [n][2] */
  int32_t* elapsed;
  float* ep_ret;   /* RecordEpisodeStatistics float32 accumulator */
  int32_t* ep_len;
  /* wrappers */
  double* o_mean; double* o_var; double* o_count;   /* [n][3], [n][3], [n] */
  double* r_mean; double* r_var; double* r_count; double* r_ret; /* [n] each */
} orc_vec_t;

void* orc_vec_create(int kind, int64_t n, int wrappers, int trig_mode, double gamma) {
  orc_vec_t* v = (orc_vec_t*)calloc(1, sizeof(orc_vec_t));
  v->kind = kind; v->n = n; v->wrappers = wrappers; v->trig_mode = trig_mode; v->gamma = gamma;
  v->rng = (pcg64_t*)calloc(n, sizeof(pcg64_t));
  v->phys = (double*)calloc(n * 4, sizeof(double));
  v->elapsed = (int32_t*)calloc(n, sizeof(int32_t));
  v->ep_ret = (float*)calloc(n, sizeof(float));
  v->ep_len = (int32_t*)calloc(n, sizeof(int32_t));
  v->o_mean = (double*)calloc(n * 3, sizeof(double));
  v->o_var = (double*)calloc(n * 3, sizeof(double));
  v->o_count = (double*)calloc(n, sizeof(double));
  v->r_mean = (double*)calloc(n, sizeof(double));
  v->r_var = (double*)calloc(n, sizeof(double));
  v->r_count = (double*)calloc(n, sizeof(double));
  v->r_ret = (double*)calloc(n, sizeof(double));
  for (int64_t i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) v->o_var[i * 3 + k] = 1.0;
    v->o_count[i] = 1e-4; v->r_var[i] = 1.0; v->r_count[i] = 1e-4;
  }
  return v;
}
void orc_vec_destroy(void* h) {
  orc_vec_t* v = (orc_vec_t*)h;
  free(v->rng); free(v->phys); free(v->elapsed); free(v->ep_ret); free(v->ep_len);
  free(v->o_mean); free(v->o_var); free(v->o_count);
  free(v->r_mean); free(v->r_var); free(v->r_count); free(v->r_ret); free(v);
}

/* observation of env i through NormalizeObservation + clip, cast to float32 */
static void emit_obs(orc_vec_t* v, int64_t i, const double* raw32_as_double, int d, float* out) {
  if (!v->wrappers) { for (int k = 0; k < d; ++k) out[k] = (float)raw32_as_double[k]; return; }
  for (int k = 0; k < d; ++k) {
    /* one shared count: RunningMeanStd keeps a scalar count for the whole vector */
    double cnt = v->o_count[i];
    rms_update1(&v->o_mean[i * 3 + k], &v->o_var[i * 3 + k], &cnt, raw32_as_double[k]);
    if (k == d - 1) v->o_count[i] = cnt;
  }
  for (int k = 0; k < d; ++k) {
    double z = (raw32_as_double[k] - v->o_mean[i * 3 + k]) / sqrt(v->o_var[i * 3 + k] + 1e-8);
    z = z < -10.0 ? -10.0 : (z > 10.0 ? 10.0 : z);
    out[k] = (float)z;
  }
}

static void env_reset(orc_vec_t* v, int64_t i, float* obs_out) {
  double* st = &v->phys[i * 4];
  v->elapsed[i] = 0; v->ep_ret[i] = 0.0f; v->ep_len[i] = 0;
  if (v->kind == 0) {
    for (int k = 0; k < 4; ++k) st[k] = pcg64_uniform(&v->rng[i], -0.05, 0.05 - (-0.05));
    double raw[4]; for (int k = 0; k < 4; ++k) raw[k] = (double)(float)st[k];
    emit_obs(v, i, raw, 4, obs_out);
  } else if (v->kind == 1) {
    st[0] = pcg64_uniform(&v->rng[i], -M_PI, M_PI - (-M_PI));
    st[1] = pcg64_uniform(&v->rng[i], -1.0, 1.0 - (-1.0));
    double s, c; trig(v->trig_mode, st[0], &s, &c);
    double raw[3] = {(double)(float)c, (double)(float)s, (double)(float)st[1]};
    emit_obs(v, i, raw, 3, obs_out);
  } else if (v->kind == 3) {
    /* Acrobot-v1 reset: uniform(-0.1, 0.1, size=4).astype(np.float32); the observation's sin / cos are evaluated in
     * fp64 and rounded (NumPy evaluates them in float32 on the float32 state: may differ by one ulp in rare cases) */
    for (int k = 0; k < 4; ++k) st[k] = (double)(float)pcg64_uniform(&v->rng[i], -0.1, 0.1 - (-0.1));
    double raw[6];
    acrobot_raw_obs(v->trig_mode, st, raw);
    emit_obs(v, i, raw, 6, obs_out);
  } else {
    /* MountainCar-v0 / MountainCarContinuous-v0 (reset): state = [uniform(-0.6, -0.4), 0] */
    st[0] = pcg64_uniform(&v->rng[i], -0.6, -0.4 - (-0.6));
    st[1] = 0.0;
    double raw[2] = {(double)(float)st[0], (double)(float)st[1]};
    emit_obs(v, i, raw, 2, obs_out);
  }
}

/* seeds: PCG64 (state, inc) pairs already expanded by NumPy's SeedSequence on the
 * Python side: [n][4] = state_hi, state_lo, inc_hi, inc_lo */
void orc_vec_reset(void* h, const uint64_t* pcg, float* obs_out) {
  orc_vec_t* v = (orc_vec_t*)h;
  int d = obs_dim_of(v->kind);
  for (int64_t i = 0; i < v->n; ++i) {
    v->rng[i].state = ((u128)pcg[i * 4 + 0] << 64) | pcg[i * 4 + 1];
    v->rng[i].inc = ((u128)pcg[i * 4 + 2] << 64) | pcg[i * 4 + 3];
    env_reset(v, i, &obs_out[i * d]);
  }
}

/* One SyncVectorEnv.step.  actions: CartPole / MountainCar int32 [n]; Pendulum float32 [n].
 * Outputs: obs [n][d] float32 (the RESET obs where an episode ended), reward
 * [n] float64, terminated/truncated [n] uint8, and for finished episodes
 * final_ret [n] float32 / final_len [n] int32 (else len = 0). */
void orc_vec_step(void* h, const void* actions, float* obs_out, double* rew_out,
                  uint8_t* term_out, uint8_t* trunc_out, float* final_ret, int32_t* final_len) {
  orc_vec_t* v = (orc_vec_t*)h;
  for (int64_t i = 0; i < v->n; ++i) {
    double* st = &v->phys[i * 4];
    double reward; int terminated = 0, truncated = 0;
    float* o = &obs_out[i * obs_dim_of(v->kind)];
    double raw[6]; int d;
    if (v->kind == 3) {
      terminated = acrobot_step(v->trig_mode, st, ((const int32_t*)actions)[i]);
      reward = terminated ? 0.0 : -1.0;
      acrobot_raw_obs(v->trig_mode, st, raw);
      d = 6;
    } else if (v->kind == 4) {
      /* MountainCarContinuous-v0 (continuous_mountain_car.py step): float32 state array, float64 arithmetic */
      const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.45, goal_velocity = 0.0;
      const double power = 0.0015;
      float a32 = ((const float*)actions)[i];
      if (v->wrappers) a32 = a32 < -1.0f ? -1.0f : (a32 > 1.0f ? 1.0f : a32);     /* ClipAction: Box(-1, 1) */
      float f32 = a32 < -1.0f ? -1.0f : (a32 > 1.0f ? 1.0f : a32);                /* min(max(action[0], -1), 1) */
      double position = st[0], velocity = st[1];
      double s3, c3; trig(v->trig_mode, 3 * position, &s3, &c3);
      velocity = velocity + ((double)f32 * power - 0.0025 * c3);
      if (velocity > max_speed) velocity = max_speed;
      if (velocity < -max_speed) velocity = -max_speed;
      position = position + velocity;
      if (position > max_position) position = max_position;
      if (position < min_position) position = min_position;
      if (position == min_position && velocity < 0) velocity = 0;
      terminated = (position >= goal_position) && (velocity >= goal_velocity);
      reward = 0;
      if (terminated) reward = 100.0;
      reward -= ((double)a32 * (double)a32) * 0.1;                                /* math.pow(action[0], 2) * 0.1 */
      st[0] = (double)(float)position; st[1] = (double)(float)velocity;           /* np.array([...], dtype=np.float32) */
      raw[0] = st[0]; raw[1] = st[1];
      d = 2;
    } else if (v->kind == 0) {
      const double gravity = 9.8, masscart = 1.0, masspole = 0.1, length = 0.5, force_mag = 10.0, tau = 0.02;
      const double total_mass = masspole + masscart, polemass_length = masspole * length;
      const double theta_thr = 12 * 2 * M_PI / 360, x_thr = 2.4;
      int a = ((const int32_t*)actions)[i];
      double x = st[0], x_dot = st[1], theta = st[2], theta_dot = st[3];
      double force = a == 1 ? force_mag : -force_mag;
      double sintheta, costheta; trig(v->trig_mode, theta, &sintheta, &costheta);
      double temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass;
      double thetaacc = (gravity * sintheta - costheta * temp) /
                        (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
      double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
      x = x + tau * x_dot; x_dot = x_dot + tau * xacc;
      theta = theta + tau * theta_dot; theta_dot = theta_dot + tau * thetaacc;
      st[0] = x; st[1] = x_dot; st[2] = theta; st[3] = theta_dot;
      terminated = (x < -x_thr) || (x > x_thr) || (theta < -theta_thr) || (theta > theta_thr);
      reward = 1.0;
      for (int k = 0; k < 4; ++k) raw[k] = (double)(float)st[k];
      d = 4;
    } else if (v->kind == 2) {
      /* MountainCar-v0 (gym/envs/classic_control/mountain_car.py step) */
      const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.5, goal_velocity = 0.0;
      const double force = 0.001, gravity = 0.0025;
      int a = ((const int32_t*)actions)[i];
      double position = st[0], velocity = st[1];
      double s3, c3; trig(v->trig_mode, 3 * position, &s3, &c3);
      velocity = velocity + ((a - 1) * force + c3 * (-gravity));
      velocity = velocity < -max_speed ? -max_speed : (velocity > max_speed ? max_speed : velocity);   /* np.clip */
      position = position + velocity;
      position = position < min_position ? min_position : (position > max_position ? max_position : position);
      if (position == min_position && velocity < 0) velocity = 0;
      terminated = (position >= goal_position) && (velocity >= goal_velocity);
      reward = -1.0;
      st[0] = position; st[1] = velocity;
      raw[0] = (double)(float)position; raw[1] = (double)(float)velocity;
      d = 2;
    } else {
      const double max_speed = 8.0, max_torque = 2.0, dt = 0.05, g = 10.0, m = 1.0, l = 1.0;
      float u32 = ((const float*)actions)[i];
      if (v->wrappers) u32 = u32 < -2.0f ? -2.0f : (u32 > 2.0f ? 2.0f : u32);   /* ClipAction */
      u32 = u32 < (float)-max_torque ? (float)-max_torque : (u32 > (float)max_torque ? (float)max_torque : u32);
      float usq32 = u32 * u32;
      double th = st[0], thdot = st[1];
      double twopi = 2 * M_PI;
      double an = fmod(th + M_PI, twopi);
      if (an != 0.0 && an < 0.0) an += twopi;    /* Python float % */
      an = an - M_PI;
      double costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (double)usq32;
      double s, c; trig(v->trig_mode, th, &s, &c);
      double newthdot = thdot + (3 * g / (2 * l) * s + 3.0 / (m * (l * l)) * (double)u32) * dt;
      newthdot = newthdot < -max_speed ? -max_speed : (newthdot > max_speed ? max_speed : newthdot);
      double newth = th + newthdot * dt;
      st[0] = newth; st[1] = newthdot;
      reward = -costs;
      trig(v->trig_mode, newth, &s, &c);
      raw[0] = (double)(float)c; raw[1] = (double)(float)s; raw[2] = (double)(float)newthdot;
      d = 3;
    }
    /* TimeLimit */
    v->elapsed[i] += 1;
    if (v->elapsed[i] >= time_limit_of(v->kind)) truncated = 1;
    /* RecordEpisodeStatistics (raw reward, float32 accumulator) */
    v->ep_ret[i] = v->ep_ret[i] + (float)reward;
    v->ep_len[i] += 1;
    final_len[i] = 0; final_ret[i] = 0.0f;
    if (terminated || truncated) { final_ret[i] = v->ep_ret[i]; final_len[i] = v->ep_len[i]; }
    /* NormalizeObservation + clip for the stepped obs (updates the running stats
     * even when the env is about to be reset: the wrapper runs before autoreset) */
    float stepped[6];
    emit_obs(v, i, raw, d, stepped);
    if (v->wrappers) {
      /* NormalizeReward + clip */
      v->r_ret[i] = v->r_ret[i] * v->gamma + reward;
      rms_update1(&v->r_mean[i], &v->r_var[i], &v->r_count[i], v->r_ret[i]);
      reward = reward / sqrt(v->r_var[i] + 1e-8);
      if (terminated || truncated) v->r_ret[i] = 0.0;
      reward = reward < -10.0 ? -10.0 : (reward > 10.0 ? 10.0 : reward);
    }
    rew_out[i] = reward; term_out[i] = (uint8_t)terminated; trunc_out[i] = (uint8_t)truncated;
    if (terminated || truncated) env_reset(v, i, o);     /* SyncVectorEnv autoreset */
    else for (int k = 0; k < d; ++k) o[k] = stepped[k];
  }
}

/* raw fp64 physical state, [n][S] (S = 4 CartPole, 2 Pendulum) */
void orc_vec_get_phys(void* h, double* out) {
  orc_vec_t* v = (orc_vec_t*)h; int S = (v->kind == 0 || v->kind == 3) ? 4 : 2;
  for (int64_t i = 0; i < v->n; ++i) for (int k = 0; k < S; ++k) out[i * S + k] = v->phys[i * 4 + k];
}
/* wrapper statistics, [n][2 D + 5] (D = obs dim; Pendulum 11): o_mean[D], o_var[D], o_count, r_mean, r_var, r_count, r_ret */
void orc_vec_get_norm(void* h, double* out) {
  orc_vec_t* v = (orc_vec_t*)h;
  int D = obs_dim_of(v->kind);
  for (int64_t i = 0; i < v->n; ++i) {
    double* o = &out[i * (2 * D + 5)];
    for (int k = 0; k < D; ++k) { o[k] = v->o_mean[i * 3 + k]; o[D + k] = v->o_var[i * 3 + k]; }
    o[2 * D] = v->o_count[i]; o[2 * D + 1] = v->r_mean[i]; o[2 * D + 2] = v->r_var[i]; o[2 * D + 3] = v->r_count[i];
    o[2 * D + 4] = v->r_ret[i];
  }
}
