// Action sink: the action half of pre_physics_step (clamp of vec_env_rlgames.py:60, current_actions stash QPC:204, take_action
// robot.py:444-461) as a device function, so that the kernel that PRODUCES the actions (the MLP head epilogue in evaluation,
// the rollout writer's sampler in training) delivers them straight into the task's buffers.  The next lm_task_pre_step is then
// called with actions == NULL and runs only the reset half: one launch, one 48-byte read and one dependent kernel boundary
// less per env step.  Arithmetic is k_pre's, operation for operation (bit-identical targets).
#pragma once
#include "lm_common.cuh"

namespace lm {

// One env's 12 actions (three float4 chunks q = 0..2 of global env e).
// do_ring / do_targets: which of the two destinations this caller writes (callers may split them over warps).
__device__ __forceinline__ void sink_store(const LmActionSink& s, int e, int q, float4 v, bool do_ring = true, bool do_targets = true) {
  if (s.clamp_actions) {
    v.x = fminf(fmaxf(v.x, -s.clip_actions), s.clip_actions); v.y = fminf(fmaxf(v.y, -s.clip_actions), s.clip_actions);
    v.z = fminf(fmaxf(v.z, -s.clip_actions), s.clip_actions); v.w = fminf(fmaxf(v.w, -s.clip_actions), s.clip_actions);
  }
  if (do_ring) reinterpret_cast<float4*>(s.ring_next + (size_t)e * LM_NUM_ACT)[q] = v;
  float* tg = e < s.loco_envs ? s.targets_loco : s.targets_mani;
  if (tg == nullptr || !do_targets) return;
  const int hr = e < s.loco_envs ? e : e - s.loco_envs;
  if (s.zero_actions) v = make_float4(0.f, 0.f, 0.f, 0.f);
  const float range = sub(s.action_limit, -s.action_limit);
  const float offset = mul(add(-s.action_limit, s.action_limit), 0.5f);
  float4 t;
  t.x = add(mul(mul(v.x, range), 0.5f), offset); t.y = add(mul(mul(v.y, range), 0.5f), offset);
  t.z = add(mul(mul(v.z, range), 0.5f), offset); t.w = add(mul(mul(v.w, range), 0.5f), offset);
  reinterpret_cast<float4*>(tg + (size_t)hr * LM_NUM_ACT)[q] = t;
}

}  // namespace lm
