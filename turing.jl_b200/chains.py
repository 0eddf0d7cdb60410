"""Chain container with the observable contract of `MCMCChains.Chains` / `VNChain`
(SURVEY.md section 8b): array (iterations x names x chains); parameter columns in model order;
internals `logprior, loglikelihood, logjoint` in user space followed by the sampler stats."""
import numpy as np

from ._lib import STAT_NAMES

INTERNALS = ("logprior", "loglikelihood", "logjoint") + STAT_NAMES
#: columns AdvancedHMC also reports; derived on access instead of stored
DERIVED = ("is_accept", "nom_step_size")


class Chains:
    def __init__(self, value, param_names, internal_names=INTERNALS, info=None):
        self.value = value  # [iters, names, chains]
        self.param_names = tuple(param_names)
        self.internal_names = tuple(internal_names)
        self.names = self.param_names + self.internal_names
        self.info = info or {}
        assert value.shape[1] == len(self.names)

    @property
    def shape(self):
        return self.value.shape

    def size(self):
        return self.value.shape

    def __len__(self):
        return self.value.shape[0]

    def __getitem__(self, name):
        """chain[name] -> [iters, chains]"""
        if isinstance(name, str):
            if name in ("nom_step_size", "SGLD_stepsize"):   # src/mcmc/sghmc.jl:238
                return self["step_size"]
            if name == "is_accept":
                # NUTS always accepts; HMC/HMCDA carry their accept flag in the tree_depth column
                if self.info.get("algorithm", 0) == 0:
                    return np.where(np.isnan(self["n_steps"]), np.nan, 1.0)
                return self["tree_depth"]
            return self.value[:, self.names.index(name), :]
        raise KeyError(name)

    def get_params(self):
        return self.value[:, :len(self.param_names), :]

    def mean(self, name=None):
        if name is not None:
            return float(np.nanmean(self[name]))
        return {n: float(np.nanmean(self[n])) for n in self.param_names}

    def var(self, name):
        return float(np.nanvar(self[name], ddof=1))

    @property
    def sampling_time(self):
        """per-chain wall time (test/mcmc/chains.jl:291-306): all chains share one launch"""
        return self.info.get("sampling_time")

    def ess_bulk(self):
        """device-computed bulk ESS per parameter (sample(..., diagnostics=True))"""
        return self.info.get("ess_bulk")

    def rhat(self):
        return self.info.get("rhat")

    def __repr__(self):
        it, _, ch = self.value.shape
        return "Chains(iterations=%d, chains=%d, parameters=%s)" % (it, ch, list(self.param_names))


def chainscat(chains):
    """concatenate along the chain dimension (AbstractMCMC.chainsstack)"""
    v = np.concatenate([c.value for c in chains], axis=2)
    info = dict(chains[0].info)
    return Chains(v, chains[0].param_names, chains[0].internal_names, info)


class ParamsWithStats:
    """One transition in user space, the counterpart of `DynamicPPL.ParamsWithStats` as the
    reference's samplers return it from `AbstractMCMC.step` (src/mcmc/hmc.jl:195-205, 258-263):
    `.params` {name: constrained value}, `.stats` {logprior, loglikelihood, logjoint, sampler stats}."""

    def __init__(self, params, stats):
        self.params = params
        self.stats = stats

    def __repr__(self):
        return "ParamsWithStats(params=%r, stats=%r)" % (self.params, self.stats)


def params_with_stats(model, sampler, transition, state=None, *, params=True, stats=False, extras=False):
    """`AbstractMCMC.ParamsWithStats(model, sampler, transition, state; params, stats, extras)` as
    overloaded in src/mcmc/Inference.jl:88-101: parameters as (string name, value) pairs taken from
    the TRANSITION (user space), not from the sampler state (linked space); stats only on request;
    no extras.  Returns (params or None, stats dict, extras dict)."""
    p = [(str(k), v) for k, v in transition.params.items()] if params else None
    s = dict(transition.stats) if stats else {}
    return p, s, {}


def _base_and_index(name):
    """'θ̃[3]' -> ('θ̃', (3,));  'x[2,1]' -> ('x', (2, 1));  'μ' -> ('μ', ())"""
    if name.endswith("]") and "[" in name:
        base, idx = name[:-1].split("[", 1)
        return base, tuple(int(t) for t in idx.split(","))
    return name, ()


class VNChain:
    """Chain keyed by variable, the observable contract of the reference's default chain type
    (`const DEFAULT_CHAIN_TYPE = VNChain`, src/mcmc/Inference.jl:65; a FlexiChains.FlexiChain keyed by
    VarName).  `chain["θ̃"]` is the whole variable, an array (iterations, chains, *shape) assembled
    from the flat columns `θ̃[1]` ... `θ̃[J]`; `chain["θ̃[2]"]` one element (iterations, chains);
    the non-parameter columns (`logjoint`, `n_steps`, ...) are the chain's extras.  The sampler
    state saved with `save_state=true` is `last_sampler_state` (Turing.loadstate,
    src/mcmc/Inference.jl:75-77)."""

    def __init__(self, value, param_names, internal_names=INTERNALS, info=None):
        self._flat = Chains(value, param_names, internal_names, info)
        self.info = self._flat.info
        groups = {}
        for col, n in enumerate(param_names):
            base, idx = _base_and_index(n)
            groups.setdefault(base, []).append((idx, col))
        self._vars = {}
        for base, items in groups.items():
            if len(items) == 1 and items[0][0] == ():
                self._vars[base] = (None, [items[0][1]])
            else:
                nd = len(items[0][0])
                shape = tuple(max(i[0][k] for i in items) for k in range(nd))
                self._vars[base] = (shape, items)
        self.parameters = tuple(self._vars)
        self.extras = tuple(internal_names) + DERIVED

    # -- the flat view, for code written against MCMCChains.Chains --
    @property
    def value(self):
        return self._flat.value

    @property
    def names(self):
        return self._flat.names

    @property
    def param_names(self):
        return self._flat.param_names

    def to_chains(self):
        return self._flat

    def keys(self):
        return self.parameters + self.extras

    def __contains__(self, name):
        return name in self._vars or name in self._flat.names or name in DERIVED

    def __len__(self):
        return len(self._flat)

    def niters(self):
        return self._flat.value.shape[0]

    def nchains(self):
        return self._flat.value.shape[2]

    def __getitem__(self, name):
        if name in self._vars:
            shape, items = self._vars[name]
            v = self._flat.value
            if shape is None:
                return v[:, items[0], :]
            out = np.full((v.shape[0], v.shape[2]) + shape, np.nan)
            for idx, col in items:
                out[(slice(None), slice(None)) + tuple(i - 1 for i in idx)] = v[:, col, :]
            return out
        return self._flat[name]

    def mean(self, name=None):
        if name is None:
            return {n: np.nanmean(self[n], axis=(0, 1)) for n in self.parameters}
        return np.nanmean(self[name], axis=(0, 1))

    @property
    def last_sampler_state(self):
        return self.info.get("samplerstate")

    @property
    def sampling_time(self):
        return self.info.get("sampling_time")

    def __repr__(self):
        it, _, ch = self._flat.value.shape
        return "VNChain(iterations=%d, chains=%d, parameters=%s)" % (it, ch, list(self.parameters))
