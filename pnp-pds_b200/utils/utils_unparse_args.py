"""tuple -> dict (utils/utils_unparse_args.py:1-36)."""
from .utils_parse_args import _CONFIGS, _EXP, _METHOD


def unparse_args_exp(gaussian_nl, sp_nl, poisson_noise, poisson_alpha, deg_op, r):
    return dict(zip((k for k, _ in _EXP), (gaussian_nl, sp_nl, poisson_noise, poisson_alpha, deg_op, r)))


def unparse_args_method(method, architecture, max_iter, gamma1, gamma2, alpha_n, alpha_s, myLambda, m1, m2, gammaInADMMStep1):
    return dict(zip((k for k, _ in _METHOD),
                    (method, architecture, max_iter, gamma1, gamma2, alpha_n, alpha_s, myLambda, m1, m2, gammaInADMMStep1)))


def unparse_args_configs(ch, add_timestamp, result_output):
    return dict(zip((k for k, _ in _CONFIGS), (ch, add_timestamp, result_output)))
