// Batched classic-control environments: one thread per env instance.
//
// Replaces the per-actor python loop body `env.step(action)` of the reference
// (jorldy/core/env/gym_env.py:61-83 Cartpole, :86-95 Pendulum/MountainCar, :32-56
// _Gym.reset/step) plus the `state = next_state if not done else env.reset()` line of
// jorldy/run_mode.py:91 / jorldy/manager/distributed_manager.py:76-92 for thousands
// of envs in one launch.  The physics restates gym 0.23.0 classic_control (third-party,
// not vendored in the reference: requirements.txt:2) from its published equations.
//
// Arithmetic: gym keeps the physical state as python floats (f64) between steps and
// returns an f32 copy as the observation; we do the same (phys[n,4] f64, obs f32) and
// use explicit round-to-nearest intrinsics so the compiler cannot contract a*b+c into
// an FMA (numpy/CPython never do), keeping the step bit-comparable with the CPU oracle
// up to the sin/cos implementation (<= 2 ulp in CUDA libm).
#include "common.cuh"
#include "philox.cuh"

namespace {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// ---- reset draws -----------------------------------------------------------------
// CartPole: U(-0.05, 0.05)^4   (gym cartpole.py reset: np_random.uniform(-0.05,0.05,(4,)))
__device__ __forceinline__ void cartpole_reset_draw(uint64_t seed, uint64_t stream, uint64_t episode,
                                                    double* s) {
  jb_philox4 a = jb_philox(seed, stream, 2 * episode);
  jb_philox4 b = jb_philox(seed, stream, 2 * episode + 1);
  double u0 = jb_u01_double(a.x, a.y), u1 = jb_u01_double(a.z, a.w);
  double u2 = jb_u01_double(b.x, b.y), u3 = jb_u01_double(b.z, b.w);
  // low + (high-low)*u  with low=-0.05, high=0.05 (numpy uniform formula)
  s[0] = dadd(-0.05, dmul(0.1, u0));
  s[1] = dadd(-0.05, dmul(0.1, u1));
  s[2] = dadd(-0.05, dmul(0.1, u2));
  s[3] = dadd(-0.05, dmul(0.1, u3));
}

struct StepOut { float reward; float done; };

__global__ void cartpole_reset_kernel(double* __restrict__ phys, float* __restrict__ obs,
                                      int32_t* __restrict__ elapsed, int64_t* __restrict__ episode,
                                      float* __restrict__ score, const uint8_t* __restrict__ mask,
                                      uint64_t seed, uint64_t stream_base, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && !mask[i]) return;
  double s[4];
  int64_t ep = episode[i];
  cartpole_reset_draw(seed, stream_base + (uint64_t)i, (uint64_t)ep, s);
  episode[i] = ep + 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) { phys[4 * i + k] = s[k]; obs[4 * i + k] = (float)s[k]; }
  elapsed[i] = 0;
  score[i] = 0.f;
}

// action_kind: 0 = int64 discrete, 1 = int32 discrete, 2 = float32 continuous (a<0 -> 0 else 1,
// gym_env.py:74-75).
__device__ __forceinline__ int read_binary_action(const void* action, int kind, int i) {
  if (kind == 0) return (int)((const int64_t*)action)[i];
  if (kind == 1) return ((const int32_t*)action)[i];
  float a = ((const float*)action)[i];
  return a < 0.f ? 0 : 1;
}

__global__ void cartpole_step_kernel(double* __restrict__ phys, float* __restrict__ obs,
                                     int32_t* __restrict__ elapsed, int64_t* __restrict__ episode,
                                     float* __restrict__ score, const void* __restrict__ action,
                                     int action_kind, float* __restrict__ next_obs,
                                     float* __restrict__ reward, float* __restrict__ done,
                                     float* __restrict__ stats /* [2]: episodes finished, sum of scores */,
                                     int auto_reset, int max_steps, uint64_t seed, uint64_t stream_base,
                                     int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double gravity = 9.8, masspole = 0.1, total_mass = 1.1 /* masspole + masscart */;
  const double length = 0.5, polemass_length = 0.05 /* masspole*length */, force_mag = 10.0, tau = 0.02;
  const double theta_thr = 12 * 2 * 3.141592653589793 / 360, x_thr = 2.4;

  double x = phys[4 * i + 0], x_dot = phys[4 * i + 1], theta = phys[4 * i + 2], theta_dot = phys[4 * i + 3];
  int a = read_binary_action(action, action_kind, i);
  double force = (a == 1) ? force_mag : -force_mag;
  double costheta = cos(theta), sintheta = sin(theta);
  // temp = (force + polemass_length * theta_dot**2 * sintheta) / total_mass
  double temp = ddiv(dadd(force, dmul(dmul(polemass_length, dmul(theta_dot, theta_dot)), sintheta)), total_mass);
  // thetaacc = (g*sin - cos*temp) / (length * (4/3 - masspole*cos**2/total_mass))
  double num = dsub(dmul(gravity, sintheta), dmul(costheta, temp));
  double den = dmul(length, dsub(4.0 / 3.0, ddiv(dmul(masspole, dmul(costheta, costheta)), total_mass)));
  double thetaacc = ddiv(num, den);
  // xacc = temp - polemass_length*thetaacc*cos/total_mass
  double xacc = dsub(temp, ddiv(dmul(dmul(polemass_length, thetaacc), costheta), total_mass));
  // explicit euler
  x = dadd(x, dmul(tau, x_dot));
  x_dot = dadd(x_dot, dmul(tau, xacc));
  theta = dadd(theta, dmul(tau, theta_dot));
  theta_dot = dadd(theta_dot, dmul(tau, thetaacc));

  bool term = (x < -x_thr) || (x > x_thr) || (theta < -theta_thr) || (theta > theta_thr);
  int el = elapsed[i] + 1;
  bool d = term || (el >= max_steps);           // gym TimeLimit: truncation also sets done
  float sc = score[i] + 1.0f;                   // env.score accumulates the gym reward (+1/step)

  next_obs[4 * i + 0] = (float)x; next_obs[4 * i + 1] = (float)x_dot;
  next_obs[4 * i + 2] = (float)theta; next_obs[4 * i + 3] = (float)theta_dot;
  reward[i] = d ? -1.0f : 0.1f;                 // JORLDY override, gym_env.py:78
  done[i] = d ? 1.0f : 0.0f;

  if (d && auto_reset) {
    if (stats) { atomicAdd(&stats[0], 1.0f); atomicAdd(&stats[1], sc); }
    double s[4];
    int64_t ep = episode[i];
    cartpole_reset_draw(seed, stream_base + (uint64_t)i, (uint64_t)ep, s);
    episode[i] = ep + 1;
    x = s[0]; x_dot = s[1]; theta = s[2]; theta_dot = s[3];
    el = 0; sc = 0.f;
  }
  phys[4 * i + 0] = x; phys[4 * i + 1] = x_dot; phys[4 * i + 2] = theta; phys[4 * i + 3] = theta_dot;
  obs[4 * i + 0] = (float)x; obs[4 * i + 1] = (float)x_dot; obs[4 * i + 2] = (float)theta; obs[4 * i + 3] = (float)theta_dot;
  elapsed[i] = el;
  score[i] = sc;
}

// ---- Pendulum-v1 (gym 0.23 pendulum.py): obs (cos th, sin th, thdot), action in [-1,1] rescaled
// to [-2,2] by _Gym.step (gym_env.py:41-45); 200-step TimeLimit; never terminates otherwise.
__device__ __forceinline__ double angle_normalize(double x) {
  // ((x + pi) % (2*pi)) - pi with python's floored modulo
  const double pi = 3.141592653589793, two_pi = 2 * 3.141592653589793;
  double y = dadd(x, pi);
  double m = fmod(y, two_pi);
  if (m < 0) m = dadd(m, two_pi);
  return dsub(m, pi);
}

__device__ __forceinline__ void pendulum_reset_draw(uint64_t seed, uint64_t stream, uint64_t episode, double* s) {
  jb_philox4 a = jb_philox(seed, stream, 2 * episode);
  const double pi = 3.141592653589793;
  double u0 = jb_u01_double(a.x, a.y), u1 = jb_u01_double(a.z, a.w);
  s[0] = dadd(-pi, dmul(2 * pi, u0));   // theta ~ U(-pi, pi)
  s[1] = dadd(-1.0, dmul(2.0, u1));     // thetadot ~ U(-1, 1)
}

__global__ void pendulum_reset_kernel(double* __restrict__ phys, float* __restrict__ obs,
                                      int32_t* __restrict__ elapsed, int64_t* __restrict__ episode,
                                      float* __restrict__ score, const uint8_t* __restrict__ mask,
                                      uint64_t seed, uint64_t stream_base, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && !mask[i]) return;
  double s[2];
  int64_t ep = episode[i];
  pendulum_reset_draw(seed, stream_base + (uint64_t)i, (uint64_t)ep, s);
  episode[i] = ep + 1;
  phys[2 * i] = s[0]; phys[2 * i + 1] = s[1];
  obs[3 * i] = (float)cos(s[0]); obs[3 * i + 1] = (float)sin(s[0]); obs[3 * i + 2] = (float)s[1];
  elapsed[i] = 0; score[i] = 0.f;
}

__global__ void pendulum_step_kernel(double* __restrict__ phys, float* __restrict__ obs,
                                     int32_t* __restrict__ elapsed, int64_t* __restrict__ episode,
                                     float* __restrict__ score, const float* __restrict__ action,
                                     float* __restrict__ next_obs, float* __restrict__ reward,
                                     float* __restrict__ done, float* __restrict__ stats, int auto_reset,
                                     int max_steps, uint64_t seed, uint64_t stream_base, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double max_speed = 8.0, max_torque = 2.0, dt = 0.05, g = 10.0, m = 1.0, l = 1.0;
  double th = phys[2 * i], thdot = phys[2 * i + 1];
  // _Gym.step rescale (f32 numpy arithmetic, gym_env.py:41-45): ((a+1)/2)*(high-low)+low
  float af = action[i];
  float scaled = ((af + 1.0f) / 2.0f) * (2.0f - (-2.0f)) + (-2.0f);
  double u = (double)scaled;
  u = fmin(fmax(u, -max_torque), max_torque);
  double an = angle_normalize(th);
  double costs = dadd(dadd(dmul(an, an), dmul(0.1, dmul(thdot, thdot))), dmul(0.001, dmul(u, u)));
  // newthdot = thdot + (3*g/(2*l)*sin(th) + 3/(m*l**2)*u)*dt
  double acc = dadd(dmul(ddiv(dmul(3, g), dmul(2, l)), sin(th)), dmul(ddiv(3.0, dmul(m, dmul(l, l))), u));
  double newthdot = dadd(thdot, dmul(acc, dt));
  newthdot = fmin(fmax(newthdot, -max_speed), max_speed);
  double newth = dadd(th, dmul(newthdot, dt));
  int el = elapsed[i] + 1;
  bool d = el >= max_steps;
  float r = (float)(-costs);
  float sc = score[i] + r;
  next_obs[3 * i] = (float)cos(newth); next_obs[3 * i + 1] = (float)sin(newth); next_obs[3 * i + 2] = (float)newthdot;
  reward[i] = r; done[i] = d ? 1.f : 0.f;
  if (d && auto_reset) {
    if (stats) { atomicAdd(&stats[0], 1.0f); atomicAdd(&stats[1], sc); }
    double s[2];
    int64_t ep = episode[i];
    pendulum_reset_draw(seed, stream_base + (uint64_t)i, (uint64_t)ep, s);
    episode[i] = ep + 1;
    newth = s[0]; newthdot = s[1]; el = 0; sc = 0.f;
  }
  phys[2 * i] = newth; phys[2 * i + 1] = newthdot;
  obs[3 * i] = (float)cos(newth); obs[3 * i + 1] = (float)sin(newth); obs[3 * i + 2] = (float)newthdot;
  elapsed[i] = el; score[i] = sc;
}

// ---- MountainCar-v0 (gym 0.23 mountain_car.py): obs (position, velocity), 3 actions,
// reward -1 per step, done at position>=0.5 and velocity>=0; 200-step TimeLimit.
__device__ __forceinline__ void mcar_reset_draw(uint64_t seed, uint64_t stream, uint64_t episode, double* s) {
  jb_philox4 a = jb_philox(seed, stream, 2 * episode);
  double u0 = jb_u01_double(a.x, a.y);
  s[0] = dadd(-0.6, dmul(0.2, u0));   // position ~ U(-0.6, -0.4)
  s[1] = 0.0;
}

__global__ void mcar_reset_kernel(double* __restrict__ phys, float* __restrict__ obs,
                                  int32_t* __restrict__ elapsed, int64_t* __restrict__ episode,
                                  float* __restrict__ score, const uint8_t* __restrict__ mask,
                                  uint64_t seed, uint64_t stream_base, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask && !mask[i]) return;
  double s[2];
  int64_t ep = episode[i];
  mcar_reset_draw(seed, stream_base + (uint64_t)i, (uint64_t)ep, s);
  episode[i] = ep + 1;
  phys[2 * i] = s[0]; phys[2 * i + 1] = s[1];
  obs[2 * i] = (float)s[0]; obs[2 * i + 1] = (float)s[1];
  elapsed[i] = 0; score[i] = 0.f;
}

__global__ void mcar_step_kernel(double* __restrict__ phys, float* __restrict__ obs,
                                 int32_t* __restrict__ elapsed, int64_t* __restrict__ episode,
                                 float* __restrict__ score, const void* __restrict__ action, int action_kind,
                                 float* __restrict__ next_obs, float* __restrict__ reward,
                                 float* __restrict__ done, float* __restrict__ stats, int auto_reset,
                                 int max_steps, uint64_t seed, uint64_t stream_base, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double min_position = -1.2, max_position = 0.6, max_speed = 0.07, goal_position = 0.5,
               goal_velocity = 0.0, force = 0.001, gravity = 0.0025;
  double pos = phys[2 * i], vel = phys[2 * i + 1];
  int a = (action_kind == 0) ? (int)((const int64_t*)action)[i] : ((const int32_t*)action)[i];
  // velocity += (action - 1) * force + cos(3 * position) * (-gravity)
  vel = dadd(vel, dadd(dmul((double)(a - 1), force), dmul(cos(dmul(3.0, pos)), -gravity)));
  vel = fmin(fmax(vel, -max_speed), max_speed);
  pos = dadd(pos, vel);
  pos = fmin(fmax(pos, min_position), max_position);
  if (pos == min_position && vel < 0) vel = 0;
  bool term = (pos >= goal_position) && (vel >= goal_velocity);
  int el = elapsed[i] + 1;
  bool d = term || el >= max_steps;
  float sc = score[i] - 1.0f;
  next_obs[2 * i] = (float)pos; next_obs[2 * i + 1] = (float)vel;
  reward[i] = -1.0f; done[i] = d ? 1.f : 0.f;
  if (d && auto_reset) {
    if (stats) { atomicAdd(&stats[0], 1.0f); atomicAdd(&stats[1], sc); }
    double s[2];
    int64_t ep = episode[i];
    mcar_reset_draw(seed, stream_base + (uint64_t)i, (uint64_t)ep, s);
    episode[i] = ep + 1;
    pos = s[0]; vel = s[1]; el = 0; sc = 0.f;
  }
  phys[2 * i] = pos; phys[2 * i + 1] = vel;
  obs[2 * i] = (float)pos; obs[2 * i + 1] = (float)vel;
  elapsed[i] = el; score[i] = sc;
}

}  // namespace

// kind: 0 cartpole, 1 pendulum, 2 mountain_car
JB_API int jb_env_classic_reset(int kind, double* phys, float* obs, int32_t* elapsed, int64_t* episode,
                                float* score, const uint8_t* mask, uint64_t seed, uint64_t stream_base,
                                int n, void* stream) {
  if (n <= 0 || !phys || !obs || !elapsed || !episode || !score) return JB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int threads = 128, blocks = jb_div_up(n, threads);
  if (kind == 0) cartpole_reset_kernel<<<blocks, threads, 0, s>>>(phys, obs, elapsed, episode, score, mask, seed, stream_base, n);
  else if (kind == 1) pendulum_reset_kernel<<<blocks, threads, 0, s>>>(phys, obs, elapsed, episode, score, mask, seed, stream_base, n);
  else if (kind == 2) mcar_reset_kernel<<<blocks, threads, 0, s>>>(phys, obs, elapsed, episode, score, mask, seed, stream_base, n);
  else return JB_ERR_INVALID;
  return jb_check_launch();
}

JB_API int jb_env_classic_step(int kind, double* phys, float* obs, int32_t* elapsed, int64_t* episode,
                               float* score, const void* action, int action_kind, float* next_obs,
                               float* reward, float* done, float* stats, int auto_reset, int max_steps,
                               uint64_t seed, uint64_t stream_base, int n, void* stream) {
  if (n <= 0 || !phys || !obs || !action || !next_obs || !reward || !done) return JB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  int threads = 128, blocks = jb_div_up(n, threads);
  if (kind == 0) {
    if (action_kind < 0 || action_kind > 2) return JB_ERR_INVALID;
    cartpole_step_kernel<<<blocks, threads, 0, s>>>(phys, obs, elapsed, episode, score, action, action_kind,
                                                    next_obs, reward, done, stats, auto_reset, max_steps, seed, stream_base, n);
  } else if (kind == 1) {
    if (action_kind != 2) return JB_ERR_INVALID;
    pendulum_step_kernel<<<blocks, threads, 0, s>>>(phys, obs, elapsed, episode, score, (const float*)action,
                                                    next_obs, reward, done, stats, auto_reset, max_steps, seed, stream_base, n);
  } else if (kind == 2) {
    if (action_kind != 0 && action_kind != 1) return JB_ERR_INVALID;
    mcar_step_kernel<<<blocks, threads, 0, s>>>(phys, obs, elapsed, episode, score, action, action_kind,
                                                next_obs, reward, done, stats, auto_reset, max_steps, seed, stream_base, n);
  } else return JB_ERR_INVALID;
  return jb_check_launch();
}
