"""Test infrastructure: a Fibonacci-style plonky2 circuit built by hand (the CircuitBuilder is out of scope): gate
instances, a witness that satisfies every gate, copy constraints and the sigma polynomials they induce
(WirePartition::get_sigma_polys, plonky2/src/plonk/permutation_argument.rs:113-157). Used by the CPU pins and the GPU
parity test of the plonky2 quotient."""
import numpy as np

from oracle_lib import GATE_ARITHMETIC, GATE_CONSTANT, GATE_NOOP, GATE_PUBLIC_INPUT

P = 0xFFFFFFFF00000001
G = 14293326489335486720   # MULTIPLICATIVE_GROUP_GENERATOR


def root_of_unity(bits):
    import plonky2_b200.field as F   # host-side field helpers (pure Python)

    return F.primitive_root_of_unity(bits)


def rnd(rng, shape=None):
    return rng.integers(0, P, size=shape, dtype=np.uint64)


class FibonacciCircuit:
    """Row 0: PublicInputGate; row 1: ConstantGate(2) holding (F_0, 1); then ArithmeticGate rows whose operations compute
    out = m0 * m1 + addend with m0 = previous out, m1 = the constant 1, addend = the out before that (copy
    constraints); NoopGate rows pad to 2^degree_bits. Unconstrained wires carry random values."""

    def __init__(self, plonk, config, degree_bits, seed=1, arithmetic_rows=None, break_gate=False, break_copy=False):
        rng = np.random.default_rng(seed)
        n = 1 << degree_bits
        self.config, self.n = config, n
        arith = plonk.ArithmeticGate.new_from_config(config)
        num_ops = arith.num_ops
        arithmetic_rows = arithmetic_rows if arithmetic_rows is not None else n - 5
        assert 2 + arithmetic_rows <= n
        f0 = int(rnd(rng))
        instances = [(plonk.PublicInputGate(), []), (plonk.ConstantGate(2), [f0, 1])]
        instances += [(arith, [1, 1])] * arithmetic_rows
        instances += [(plonk.NoopGate(), [])] * (n - len(instances))
        self.common, self.constant_vecs = plonk.CommonCircuitData.from_gate_instances(config, instances)
        self.public_inputs_hash = [int(v) for v in rnd(rng, 4)]
        wires = rnd(rng, (config.num_wires, n))
        wires[0:4, 0] = self.public_inputs_hash
        wires[0, 1], wires[1, 1] = f0, 1
        # partition of the routed wires: sets of (row, column) that must carry one value
        sets = {"one": [(1, 1)], "f0": [(1, 0)]}
        prev, prevprev = ("f0", f0), ("one", 1)
        t = 0
        for r in range(2, 2 + arithmetic_rows):
            for k in range(num_ops):
                m0, m1, addend = prev[1], 1, prevprev[1]
                out = (m0 * m1 + addend) % P
                wires[4 * k, r], wires[4 * k + 1, r], wires[4 * k + 2, r], wires[4 * k + 3, r] = m0, m1, addend, out
                sets[prev[0]].append((r, 4 * k))
                sets["one"].append((r, 4 * k + 1))
                sets[prevprev[0]].append((r, 4 * k + 2))
                name = "out%d" % t
                sets[name] = [(r, 4 * k + 3)]
                prevprev, prev = prev, (name, out)
                t += 1
        if break_gate:     # one arithmetic output off by one: the vanishing polynomial is no longer divisible by Z_H
            wires[3, 2] = (int(wires[3, 2]) + 1) % P
        if break_copy:     # a copy constraint violated while every gate still holds
            wires[1, 0] = (int(wires[1, 0]) + 1) % P
            sets["one"].append((0, 1))
        self.wires = wires
        # get_sigma_map: the next wire of the same set, wrapping around; a wire alone in its set maps to itself
        neighbor = {}
        for members in sets.values():
            for i, w in enumerate(members):
                neighbor[w] = members[(i + 1) % len(members)]
        k_is = self.common.k_is
        omega = root_of_unity(degree_bits)
        subgroup = [1]
        for _ in range(n - 1):
            subgroup.append(subgroup[-1] * omega % P)
        sig = np.empty((config.num_routed_wires, n), dtype=np.uint64)
        for col in range(config.num_routed_wires):
            for row in range(n):
                nr, ncol = neighbor.get((row, col), (row, col))
                sig[col, row] = k_is[ncol] * subgroup[nr] % P
        self.sigmas = sig
        self.constants_sigmas = np.concatenate([np.stack(self.constant_vecs), sig])

    def oracle_circuit(self):
        """The dict tests/oracle_lib.plonk_quotient takes, from the product's CommonCircuitData."""
        cd = self.common
        kinds = {"NoopGate": GATE_NOOP, "ConstantGate": GATE_CONSTANT, "PublicInputGate": GATE_PUBLIC_INPUT,
                 "ArithmeticGate": GATE_ARITHMETIC}
        gates = []
        for i, g in enumerate(cd.gates):
            sel = cd.selectors_info.selector_indices[i]
            grp = cd.selectors_info.groups[sel]
            param = getattr(g, "num_consts", getattr(g, "num_ops", 0))
            gates.append((kinds[g.id().split(" ")[0]], param, sel, grp.start, grp.stop))
        cfg = cd.config
        return dict(num_wires=cfg.num_wires, num_routed_wires=cfg.num_routed_wires, num_constants=cd.num_constants,
                    num_challenges=cfg.num_challenges, quotient_degree_factor=cd.quotient_degree_factor,
                    num_selectors=cd.selectors_info.num_selectors(), num_partial_products=cd.num_partial_products,
                    num_gate_constraints=cd.num_gate_constraints, k_is=cd.k_is, gates=gates)

    def oracle_zs_partial_products(self, oracle, betas, gammas):
        """[plonk_z_vecs, partial_products.concat()] (plonk/prover.rs:227-232) from the oracle's restatement."""
        cfg = self.config
        zs, pps = [], []
        for beta, gamma in zip(betas, gammas):
            out = oracle.partial_products_and_zs(self.wires[:cfg.num_routed_wires], self.sigmas, self.common.k_is, beta, gamma,
                                                 self.common.quotient_degree_factor)
            zs.append(out[-1])
            pps += list(out[:-1])
        return np.stack(zs + pps)
