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
