"""dict -> tuple argument parsing with the reference's defaults (utils/utils_parse_args.py:1-39)."""

_EXP = (("gaussian_nl", 0), ("sp_nl", 0), ("poisson_noise", False), ("poisson_alpha", 300), ("deg_op", "blur"), ("r", 0.8))
_METHOD = (("method", "ours-A"), ("architecture", "DnCNN_nobn_nch_3_nlev_0.01"), ("max_iter", 10), ("gamma1", 1), ("gamma2", 1),
           ("alpha_n", 1), ("alpha_s", 1), ("myLambda", 1), ("m1", 15), ("m2", 15), ("gammaInADMMStep1", 0.1))
_CONFIGS = (("ch", 3), ("add_timestamp", True), ("result_output", False))


def _parse(args, table):
    return tuple(args.get(k, d) for k, d in table)


def parse_args_exp(args):
    """-> gaussian_nl, sp_nl, poisson_noise, poisson_alpha, deg_op, r"""
    return _parse(args, _EXP)


def parse_args_method(args):
    """-> method, architecture, max_iter, gamma1, gamma2, alpha_n, alpha_s, myLambda, m1, m2, gammaInADMMStep1"""
    return _parse(args, _METHOD)


def parse_args_configs(args):
    """-> ch, add_timestamp, result_output"""
    return _parse(args, _CONFIGS)
