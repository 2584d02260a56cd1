from .mcmc import MCMCSampler, BlockedMCMCSampler, MCMCHistory, Metropolis  # noqa: F401
