"""Race / determinism stress tests (compute-sanitizer is not available on the GPU pool, so the cross-CTA protocols are
exercised by repetition and compared bit for bit):

  * the fused critic head's last-CTA reduction and optimiser tick (misc.cuh: head_body) and every stage of the update:
    1500 graph replays, twice from the same state -> identical bits; the same for an 8-agent population (different grids)
  * the persistent kernel's device-wide barrier (persist.cuh): 400 updates inside one cooperative launch, twice
  * the chain kernels' release / acquire tq hand-over between CTAs and their last-CTA finalisation (chain.cuh): 600 updates,
    twice, plus graph == plain launches
  * the host mirror of the loss (head_body's unfenced 8-byte store): 1000 steps of train + wait_critic_loss, the mirrored
    value equals the device value at every step
"""
import pytest
import torch

from helpers import make_featured

pytestmark = pytest.mark.gpu


def _snapshot(ours):
    return {k: {n: t.detach().clone() for n, t in getattr(ours, k).state_dict().items()}
            for k in ("actor", "critic", "actor_target", "critic_target")}


def _same(a, b, what):
    for k, sd in a.items():
        for n, t in sd.items():
            assert torch.equal(t, b[k][n]), f"{what}: {k}.{n} differs (max |d| {(t - b[k][n]).abs().max().item():.3e})"


def _population(n_agents, precision):
    from oracle import td3_oracle as O
    from td3_b200.TD3_featured import TD3
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    obs, act = O.Space(17), O.Space(6)
    torch.manual_seed(3)
    pop = TD3(obs, act, n_agents=n_agents, precision=precision, seed=5, actor_widths=(400, 300), q_widths=(400, 300), lr=1e-3)
    rb = ReplayBuffer_featured(obs, act, max_size=2048, n_agents=n_agents)
    for i in range(n_agents):
        rb.add_batch(agent=i, **O.synthetic_transitions_featured(2048, 17, 6, seed=10 + i))
    return pop, rb


def _snapshot_population(pop, n_agents):
    return {f"{k}[{i}]": {n: t.detach().clone() for n, t in pop.agent_state_dict(k, i).items()}
            for k in ("actor", "critic", "actor_target", "critic_target") for i in range(n_agents)}


def _run(mode, updates, precision="tf32", n_agents=1, chunk=None):
    if n_agents > 1:
        ours, rb = _population(n_agents, precision)
    else:
        _, _, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-3, precision=precision)
    ours.exec_mode = mode
    done = 0
    while done < updates:
        n = min(chunk or updates, updates - done)
        ours.train(rb, 256, iterations=n)
        done += n
    torch.cuda.synchronize()
    return (_snapshot_population(ours, n_agents) if n_agents > 1 else _snapshot(ours)), ours


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_graph_replays_are_bit_reproducible(precision):
    a, _ = _run("graph", 1500, precision)
    b, _ = _run("graph", 1500, precision, chunk=7)          # different replay batching, same sequence of updates
    _same(a, b, f"graph x2 ({precision})")


def test_population_grids_are_bit_reproducible():
    a, _ = _run("graph", 300, n_agents=8)
    b, _ = _run("graph", 300, n_agents=8, chunk=11)
    _same(a, b, "8-agent population x2")


def test_persistent_kernel_barrier_is_bit_reproducible():
    a, _ = _run("persistent", 400)
    b, _ = _run("persistent", 400, chunk=50)
    _same(a, b, "persistent x2")


def test_chain_kernels_flags_and_finalisation_are_bit_reproducible(monkeypatch):
    monkeypatch.setenv("TD3_CHAIN", "1")
    a, agent = _run("graph", 600)
    assert agent.chain_active()
    b, _ = _run("graph", 600, chunk=13)
    c, _ = _run("launches", 600, chunk=100)
    _same(a, b, "chain graph x2")
    _same(a, c, "chain graph vs launches")


def test_host_mirror_of_the_loss_tracks_the_device_value():
    _, _, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-3, precision="tf32")
    if not ours._status_live:
        pytest.skip("host status words not live on this configuration")
    bad = 0
    for step in range(1000):
        ours.train(rb, 256)
        mirrored = ours.wait_critic_loss()
        if step % 50 == 49:                                  # a device read drains the stream: sample it
            torch.cuda.synchronize()
            dev = float(ours.last_critic_loss[0].item())
            bad += int(abs(dev - float(mirrored)) > 0.0)
    assert bad == 0
