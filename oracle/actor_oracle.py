"""TEST INFRASTRUCTURE (checker only; never imported by the product path).

numpy float64 restatement of the actor-side n-step reward pre-sum and initial priorities
(/root/reference/actor.py:74-76 `calc_nstep_reward`, :78-107 `calc_priorities`), pinned against
tests/golden/ref_actor_prio.npz (produced by the UNMODIFIED reference, oracle/make_golden.py).

What the reference does, per finished episode of E real rows + n pad rows (actor.py:173):
  * R_i = sum_{j<n} gamma^j r_{i+j} for i < E, computed in place in increasing i (later rows are still raw);
  * all nets restart from the zero state; the target nets first consume rows 0..n-1 (actor.py:86-89), then the loop
    i = 0..E-1 feeds row i to the online critic (stored action) and row i+n to target actor -> target critic:
    every net sees rows 0, 1, 2, ... exactly once;
  * for i >= burn_in: td_i = mean_A( q_i - h(R_i + gamma^n (1 - term_{i+n-1}) q'_{i+n}) )   [mean of the DIFFERENCE];
    a deque of the last `learning` td values; for i >= burn_in + learning the priority
    0.9 max + 0.1 mean of td^2 over the deque, i.e. priority k covers steps k+burn_in+1 .. k+burn_in+learning
    (one step later than the window the learner trains on - the reference's off-by-one, kept).
"""
import numpy as np

from oracle import learner_oracle as lo


def nstep_rewards(raw, n_step, gamma):
    raw = np.asarray(raw, np.float64)
    out = raw.copy()
    for i in range(len(raw) - n_step):
        out[i] = sum(raw[i + j] * gamma ** j for j in range(n_step))
    return out


def episode_priorities(critic, target_actor, target_critic, obs, act, rew, term, *, burn_in, learning, n_step,
                       gamma, eta=0.9):
    """obs [N,O], act [N,A], rew [N] (already n-step sums), term [N]; N = E + n_step.  Returns priorities [E - burn_in - learning]."""
    f = lambda a: np.asarray(a, np.float32).astype(np.float64)  # noqa: E731
    N = obs.shape[0]
    E = N - n_step
    H = np.asarray(critic["l2.weight_hh"]).shape[1]
    P = lambda sd: {k: f(v) for k, v in sd.items()}  # noqa: E731
    z = np.zeros((1, H))
    x_c = np.concatenate((f(obs[:E]), f(act[:E])), 1)[:, None, :]
    q = lo.net_forward(P(critic), x_c, z, z, critic=True)["out"][:, 0]                     # [E, A]
    a_t = lo.net_forward(P(target_actor), f(obs)[:, None, :], z, z, critic=False)["out"]    # [N, 1, A]
    x_t = np.concatenate((f(obs)[:, None, :], a_t), 2)
    q_t = lo.net_forward(P(target_critic), x_t, z, z, critic=True)["out"][:, 0]             # [N, A]
    td = np.zeros(E)
    for i in range(burn_in, E):
        y = lo.value_rescale(rew[i] + gamma ** n_step * (1.0 - term[i + n_step - 1]) * q_t[i + n_step])
        td[i] = (q[i] - y).mean()
    out = []
    for i in range(burn_in + learning, E):
        w = td[i - learning + 1:i + 1] ** 2
        out.append(eta * w.max() + (1 - eta) * w.mean())
    return np.asarray(out, np.float64)
