"""Model check of the role / hand-shake protocol of walk_stream_kernel (tpnet_b200/csrc/tpn_update.cu): producers and
chains are CTAs of ONE launch, chains wait for flags that producers set, and nothing may depend on which CTAs happen
to be resident together (an earlier two-launch design dead-locked when its launches were serialised).  The model below
restates the kernel's scheduling rules — ticket roles, producer registration, a chaining CTA's bounded wait for a
registered producer, work selection order — and runs them under adversarial residency limits and random interleavings:
every run must terminate with every block produced and every chain item finished.  (Host-side restatement of the
protocol; the kernel itself is covered by the -m gpu parity tests.)"""
import random

import pytest


class _Cta:
    def __init__(self, sim):
        self.sim = sim
        self.pc = 'start'
        self.polls = 0
        self.item = None          # (giant, next block) of the chain item in progress
        self.prod_first = self.prod_done = self.chains_done = self.chains = False

    def step(self):
        """One scheduling quantum; returns False when the CTA has exited."""
        s = self.sim
        if self.pc == 'start':
            ticket, s.ticket = s.ticket, s.ticket + 1
            self.chains = 1 <= ticket <= min(s.total_chain_items, s.grid // 2)
            self.chains_done = not self.chains
            if not self.chains:
                s.active += 1
                self.prod_first = True
                self.pc = 'pick'
            else:
                self.pc = 'wait_producer'
        elif self.pc == 'wait_producer':
            if s.active > 0:
                self.pc = 'pick'
            else:
                self.polls += 1
                if self.polls >= s.max_polls:            # nobody registered: this CTA produces first
                    s.active += 1
                    self.prod_first = True
                    self.pc = 'pick'
        elif self.pc == 'pick':
            if self.prod_first and not self.prod_done and self._take_block():
                return True
            if not self.chains_done:
                c, s.chain_ctr = s.chain_ctr, s.chain_ctr + 1
                if c < s.total_chain_items:
                    self.item = [c // s.slices, 0]
                    self.pc = 'chain'
                    return True
                self.chains_done = True
            if not self.prod_done and self._take_block():
                return True
            return False                                 # nothing left: exit
        elif self.pc == 'produce':
            s.flags[self.block] = True                   # a producer never waits for anything
            self.pc = 'pick'
        elif self.pc == 'chain':
            g, b = self.item
            if b == s.blocks_of[g]:
                s.finished_items += 1
                self.pc = 'pick'
            elif s.flags[s.first_block[g] + b]:          # otherwise: spin on the flag
                self.item[1] += 1
        return True

    def _take_block(self):
        s = self.sim
        q, s.prod_ctr = s.prod_ctr, s.prod_ctr + 1
        if q < s.total_blocks:
            self.block = q
            self.pc = 'produce'
            return True
        self.prod_done = True
        return False


class _Sim:
    def __init__(self, rng, grid, resident, blocks_of, slices, max_polls):
        self.grid, self.slices, self.max_polls = grid, slices, max_polls
        self.blocks_of = blocks_of
        self.first_block = [sum(blocks_of[:i]) for i in range(len(blocks_of))]
        self.total_blocks = sum(blocks_of)
        self.total_chain_items = len(blocks_of) * slices
        self.flags = [False] * self.total_blocks
        self.ticket = self.active = self.prod_ctr = self.chain_ctr = self.finished_items = 0
        pending = [_Cta(self) for _ in range(grid)]
        running = []
        steps = 0
        while pending or running:
            while pending and len(running) < resident:   # a CTA becomes resident only when a slot is free
                running.append(pending.pop(rng.randrange(len(pending))))
            cta = running[rng.randrange(len(running))]
            if not cta.step():
                running.remove(cta)
            steps += 1
            assert steps < 2_000_000, 'the protocol did not terminate: dead-lock'


@pytest.mark.parametrize('resident', [1, 2, 3, 7, 40, 148])
@pytest.mark.parametrize('giants', [[5], [40, 3], [1, 1, 1, 1], [17, 9, 9, 2, 1]])
def test_stream_protocol_terminates_under_any_residency(resident, giants):
    for seed in range(25):
        rng = random.Random(1000 * resident + seed)
        grid = rng.choice([2, 3, 16, 148])
        slices = rng.choice([1, 3, 42])
        sim = _Sim(rng, grid, min(resident, grid), giants, slices, max_polls=rng.choice([1, 4, 16]))
        assert all(sim.flags), 'every production block is produced'
        assert sim.finished_items == sim.total_chain_items, 'every (giant, slice) chain finishes'
        assert sim.prod_ctr >= sim.total_blocks and sim.chain_ctr >= sim.total_chain_items
