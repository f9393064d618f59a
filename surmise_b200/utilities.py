"""Sampler dispatcher with the contract of surmise.utilities.sampler (surmise/utilities.py:9-71), used by
directbayeswoodbury_B200.fit when surmise itself is not importable (e.g. on a GPU box that only has this package).
With surmise installed its own dispatcher is used and this module is not needed."""
import importlib


class sampler(object):
    def __init__(self, logpost_func, draw_func, sampler='PTLMC_B200', **sampler_options):
        self.logpost_func = logpost_func
        self.draw_func = draw_func
        self.options = sampler_options
        self.sampler_info = {}
        self.draw_samples(sampler)

    def draw_samples(self, sampler):
        try:
            self.method = importlib.import_module('surmise_b200.utilitiesmethods.' + sampler)
        except ImportError:
            raise ValueError("sampler '%s' is not available without surmise installed "
                             "(surmise_b200 ships: PTLMC_B200, LMC_B200, metropolis_hastings_B200)." % sampler)
        self.sampler_info = self.method.sampler(self.logpost_func, self.draw_func, **self.options)
        if 'theta' not in self.sampler_info.keys():
            raise ValueError('A sample from a posterior distribution is required.')
