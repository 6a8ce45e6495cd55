"""Test infrastructure: a Fibonacci-style plonky2 circuit built by hand (the CircuitBuilder is out of scope): gate
instances, a witness that satisfies every gate, copy constraints and the sigma polynomials they induce
(WirePartition::get_sigma_polys, plonky2/src/plonk/permutation_argument.rs:113-157). Used by the CPU pins and the GPU
parity test of the plonky2 quotient."""
import numpy as np

from oracle_lib import GATE_ARITHMETIC, GATE_CONSTANT, GATE_NOOP, GATE_POSEIDON, GATE_PUBLIC_INPUT

P = 0xFFFFFFFF00000001
G = 14293326489335486720   # MULTIPLICATIVE_GROUP_GENERATOR


def root_of_unity(bits):
    import plonky2_b200.field as F   # host-side field helpers (pure Python)

    return F.primitive_root_of_unity(bits)


def rnd(rng, shape=None):
    return rng.integers(0, P, size=shape, dtype=np.uint64)


class Fp:
    """A Goldilocks element as a plain Python int: lets the product's gate code (written over expression handles) run on
    numbers, for witness generation and for evaluating the vanishing polynomial at one point."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = int(v) % P

    @staticmethod
    def _c(o):
        return o.v if isinstance(o, Fp) else int(o) % P

    def __add__(self, o):
        return Fp(self.v + Fp._c(o))

    def __sub__(self, o):
        return Fp(self.v - Fp._c(o))

    def __rsub__(self, o):
        return Fp(Fp._c(o) - self.v)

    def __mul__(self, o):
        return Fp(self.v * Fp._c(o))

    __radd__, __rmul__ = __add__, __mul__

    def __int__(self):
        return self.v


class PointVars:
    """EvaluationVarsBase over numbers: the values of the constants, wires and public-input hash at one point."""

    def __init__(self, constants, wires, public_inputs_hash, prefix=0):
        self.c, self.w, self.pih, self.prefix = constants, wires, public_inputs_hash, prefix

    def local_constant(self, i):
        return Fp(self.c[self.prefix + i])

    def local_wire(self, i):
        return Fp(self.w[i])

    def public_inputs_hash(self, i):
        return Fp(self.pih[i])

    def remove_prefix(self, n):
        return PointVars(self.c, self.w, self.pih, self.prefix + n)


def poseidon_gate_witness(plonk, inputs, swap):
    """PoseidonGenerator::run_once (gates/poseidon.rs:440-540): wire index -> value for one PoseidonGate row."""
    G_ = plonk.PoseidonGate
    t = plonk.poseidon_tables()
    wires = {G_.wire_input(i): int(inputs[i]) for i in range(12)}
    wires[G_.WIRE_SWAP] = int(swap)
    for i in range(4):
        wires[G_.wire_delta(i)] = int(swap) * (int(inputs[i + 4]) - int(inputs[i])) % P
    state = [Fp(v) for v in inputs]
    if swap:
        for i in range(4):
            state[i], state[i + 4] = state[i + 4], state[i]
    round_ctr = 0
    for r in range(4):
        state = G_._constant_layer(state, round_ctr)
        if r != 0:
            for i in range(12):
                wires[G_.wire_full_sbox_0(r, i)] = int(state[i])
        state = G_._mds_layer([G_._sbox_monomial(x) for x in state])
        round_ctr += 1
    state = [state[i] + t["FAST_FIRST_RC"][i] for i in range(12)]
    state = G_._mds_partial_layer_init(state)
    for r in range(22):
        wires[G_.wire_partial_sbox(r)] = int(state[0])
        state[0] = G_._sbox_monomial(state[0])
        if r < 21:
            state[0] = state[0] + t["FAST_RC"][r]
        state = G_._mds_partial_layer_fast(state, r)
    round_ctr += 22
    for r in range(4):
        state = G_._constant_layer(state, round_ctr)
        for i in range(12):
            wires[G_.wire_full_sbox_1(r, i)] = int(state[i])
        state = G_._mds_layer([G_._sbox_monomial(x) for x in state])
        round_ctr += 1
    for i in range(12):
        wires[G_.wire_output(i)] = int(state[i])
    return wires


class FibonacciCircuit:
    """Row 0: PublicInputGate; row 1: ConstantGate(2) holding (F_0, 1); then ArithmeticGate rows whose operations compute
    out = m0 * m1 + addend with m0 = previous out, m1 = the constant 1, addend = the out before that (copy
    constraints); NoopGate rows pad to 2^degree_bits. Unconstrained wires carry random values."""

    def __init__(self, plonk, config, degree_bits, seed=1, arithmetic_rows=None, break_gate=False, break_copy=False,
                 poseidon_rows=0, break_poseidon=False):
        rng = np.random.default_rng(seed)
        n = 1 << degree_bits
        self.config, self.n = config, n
        arith = plonk.ArithmeticGate.new_from_config(config)
        num_ops = arith.num_ops
        arithmetic_rows = arithmetic_rows if arithmetic_rows is not None else n - 5 - poseidon_rows
        assert 2 + arithmetic_rows + poseidon_rows <= n
        f0 = int(rnd(rng))
        instances = [(plonk.PublicInputGate(), []), (plonk.ConstantGate(2), [f0, 1])]
        instances += [(arith, [1, 1])] * arithmetic_rows
        instances += [(plonk.PoseidonGate(), [])] * poseidon_rows
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
        # PoseidonGate rows: a hash chain -- the first four outputs of a row are copied into the first four inputs of the
        # next one (copy constraints on PoseidonGate wires), alternating the swap flag
        self.poseidon_io = []
        prev_out = None
        for q in range(poseidon_rows):
            r = 2 + arithmetic_rows + q
            inputs = [int(v) for v in rnd(rng, 12)]
            if prev_out is not None:
                inputs[:4] = prev_out[:4]
            pw = poseidon_gate_witness(plonk, inputs, q & 1)
            for k, v in pw.items():
                wires[k, r] = v
            if prev_out is not None:
                for i in range(4):
                    sets["p%d_%d" % (q, i)] = [(r - 1, plonk.PoseidonGate.wire_output(i)), (r, plonk.PoseidonGate.wire_input(i))]
            prev_out = [pw[plonk.PoseidonGate.wire_output(i)] for i in range(12)]
            self.poseidon_io.append((inputs, q & 1, prev_out))
        if break_poseidon:  # one partial-round S-box input off by one
            r = 2 + arithmetic_rows
            k = plonk.PoseidonGate.wire_partial_sbox(7)
            wires[k, r] = (int(wires[k, r]) + 1) % P
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
                 "ArithmeticGate": GATE_ARITHMETIC, "PoseidonGate": GATE_POSEIDON}
        gates = []
        for i, g in enumerate(cd.gates):
            sel = cd.selectors_info.selector_indices[i]
            grp = cd.selectors_info.groups[sel]
            param = getattr(g, "num_consts", getattr(g, "num_ops", 0))
            gates.append((kinds[g.id().split(" ")[0].split("(")[0]], param, sel, grp.start, grp.stop))
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
