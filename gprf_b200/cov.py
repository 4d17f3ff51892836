"""treegp-style covariance record (``treegp.gp.GPCov`` as used at gprf.py:163,
synthetic.py:149, run_seismic.py:299-301): a plain parameter holder."""
import numpy as np

DFN_IDS = {"euclidean": 0, "lld": 1}
WFN_IDS = {"se": 0, "matern32": 1}


class GPCov(object):
    def __init__(self, wfn_params, dfn_params, dfn_str="euclidean", wfn_str="se", **_ignored):
        self.wfn_params = np.array(wfn_params, dtype=float).reshape(-1)
        self.dfn_params = np.array(dfn_params, dtype=float).reshape(-1)
        self.dfn_str = dfn_str
        self.wfn_str = wfn_str

    def ids(self):
        try:
            return DFN_IDS[self.dfn_str], WFN_IDS[self.wfn_str]
        except KeyError:
            raise ValueError("unsupported covariance %s/%s (supported: %s x %s)"
                             % (self.dfn_str, self.wfn_str, sorted(DFN_IDS), sorted(WFN_IDS)))

    def __repr__(self):
        return "GPCov(wfn=%s%s, dfn=%s%s)" % (self.wfn_str, self.wfn_params.tolist(),
                                               self.dfn_str, self.dfn_params.tolist())
