from .mcmc import MCMCSampler, MCMCHistory, Metropolis  # noqa: F401
