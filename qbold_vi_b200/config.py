"""Two-tier configuration of the reference, reproduced as the drop-in input contract.

(1) physics INI ``config`` -> mapping of *strings* (``configparser['DEFAULT']``), cast ad hoc
    where used (reference signals.py:29-46, 255-267; train.py:189-191);
(2) hyper-parameters: ``get_defaults()`` -> argparse -> yaml overrides with the reference's
    typing rule (reference train.py:150-186, 454-480; utils.py:86-123).
"""
from __future__ import annotations

import argparse
import configparser
import os
from types import SimpleNamespace

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CONFIG_PATH = os.path.join(_HERE, 'configs', 'config')
OPTIMAL_YAML_PATH = os.path.join(os.path.dirname(_HERE), 'configurations', 'optimal.yaml')


def load_system_parameters(path=None):
    """``config.read('config'); params = config['DEFAULT']`` (train.py:189-191).
    Returns the live configparser section: values are strings, callers may mutate it in
    place (``params['simulate_noise'] = 'False'``, train.py:256)."""
    cfg = configparser.ConfigParser()
    path = path or ('config' if os.path.isfile('config') else DEFAULT_CONFIG_PATH)
    if not cfg.read(path):
        raise FileNotFoundError(path)
    return cfg['DEFAULT']


def get_defaults():
    """train.py:150-186."""
    return dict(no_units=30, no_intermediate_layers=1, student_t_df=2, pt_lr=5e-5, ft_lr=5e-3, kl_weight=1.0,
                smoothness_weight=1.0, dropout_rate=0.0, no_pt_epochs=5, no_ft_epochs=40, im_loss_sigma=0.08,
                crop_size=16, use_layer_norm=False, activation='relu', use_r2p_loss=False,
                multi_image_normalisation=True, full_model=True, use_blood=True, misalign_prob=0.0,
                use_population_prior=False, use_wandb=True, inv_gamma_alpha=0.0, inv_gamma_beta=0.0,
                gate_offset=0.0, resid_init_std=1e-1, channelwise_gating=True, infer_inv_gamma=False,
                use_mvg=False, uniform_prop=0.1, use_swa=True, adamw_decay=2e-4, pt_adamw_decay=2e-4,
                predict_log_data=True)


def setup_argparser(defaults):
    """train.py:107-147 (``type=bool`` flags: any non-empty CLI string is True, as in the reference)."""
    p = argparse.ArgumentParser(description='Train neural network for parameter estimation')
    p.add_argument('-f', default='synthetic_data.npz')
    p.add_argument('-d', default='/home/data/qbold/')
    p.add_argument('--save_directory', default=None)
    for key, val in defaults.items():
        p.add_argument('--' + key, type=type(val), default=val)
    return p


def apply_yaml_overrides(args, opt):
    """train.py:473-480: ``if args.get(key): args[key] = type(args[key])(val) else: args[key] = val``
    -- falsy defaults (0.0, False, None) take the yaml value untyped; unknown keys are added."""
    for key, val in opt.items():
        if args.get(key):
            args[key] = type(args.get(key))(val)
        else:
            args[key] = val
    return args


def load_arguments(argv=None, yaml_file=None):
    """utils.py:86-123 / train.py:454-480.  ``argv`` may be ``[<file>.yaml]`` like the reference CLI."""
    import yaml
    argv = list(argv or [])
    if yaml_file is None and len(argv) == 1 and '.yaml' in argv[0]:
        yaml_file, argv = argv[0], []
    args = vars(setup_argparser(get_defaults()).parse_args(argv))
    if yaml_file is not None:
        with open(yaml_file) as fh:
            apply_yaml_overrides(args, yaml.load(fh, Loader=yaml.FullLoader))
    return SimpleNamespace(**args)


def optimal_arguments():
    return load_arguments(yaml_file=OPTIMAL_YAML_PATH)
