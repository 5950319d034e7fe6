"""CLI entry with the reference's flags (jorldy/main.py:8-29):
    python -m jorldy_b200.main [--single|--sync|--async|--eval] --config config.dqn.cartpole [--domain.key value ...]
"""
import argparse

from .run_mode import async_distributed_train, evaluate, single_train, sync_distributed_train

default_config_path = "config.dqn.cartpole"


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--single", action="store_true", help="the reference's per-step loop, one env (default)")
    parser.add_argument("--sync", action="store_true",
                        help="GPU-resident pipeline: train.num_workers batched envs + learner on one device (one rank per GPU under torchrun)")
    parser.add_argument("--async", action="store_true",
                        help="same pipeline as --sync: collection and learning share the device, there is no separate interact process "
                             "to be asynchronous with")
    parser.add_argument("--eval", action="store_true", help="greedy episodes from train.load_path")
    parser.add_argument("--config", type=str, help="config.dqn.cartpole")
    args, unknown = parser.parse_known_args(argv)
    n_modes = args.single + args.sync + args.__dict__["async"] + args.eval
    assert n_modes < 2, "You have to choose only one mode"
    config_path = args.config if args.config else default_config_path
    if args.single or n_modes == 0:
        single_train(config_path, unknown)
    elif args.sync:
        sync_distributed_train(config_path, unknown)
    elif args.__dict__["async"]:
        async_distributed_train(config_path, unknown)
    elif args.eval:
        evaluate(config_path, unknown)


if __name__ == "__main__":
    main()
