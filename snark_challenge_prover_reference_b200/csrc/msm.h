// Host-callable entry points of msm.cu (internal; the public surface is include/b200_groth16.h).
#pragma once
#include <stddef.h>
#include <functional>
#include <string>
namespace b200 {
// Host tail of a deferred MSM (wait for the bucket reduction, fetch the window sums, serial window combine). Runs on any
// thread; returns 0 or the failing call's code and copies that call's message into `err` (last_error() is thread-local,
// so the issuing thread could not see it otherwise).
typedef std::function<int(std::string &err)> MsmTail;
// group: 1 = G1, 2 = G2. d_scalars: n Fr (Montgomery). d_points: n affine wire-format points. h_out: projective.
int msm_dispatch(int curve, int group, const void *d_scalars, const void *d_points, size_t n, void *h_out);
// GPU work done on return; `tail` finishes the result into h_out (serial host Horner) - run it on any thread.
int msm_dispatch_deferred(int curve, int group, const void *d_scalars, const void *d_points, size_t n, void *h_out,
                          MsmTail &tail);
void msm_set_window(int c);
// MSM streams created by the calling thread from now on get the highest priority (worker threads of prove_batch)
void msm_thread_high_priority(bool on);
// bucket accumulation: 0 = XYZZ mixed additions per task (default), 1 = rounds of batched affine additions
void msm_set_batch_affine(int on);
void msm_last_phase_ms(double *out5);
// window width c, number of windows W and task length T of the most recently prepared MSM
void msm_last_plan(int *out3);
// accumulated phase times since the last reset: out10 = G1 {digits, sort, accumulate, reduce, host}, then G2
void msm_phase_totals(double *out10, int reset);
void msm_release_workspace();
}  // namespace b200
