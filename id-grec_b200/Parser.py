"""Command line of the reference entry point (Parser.py:4-17): same five flags, same defaults."""
import argparse


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="ID-GRec (B200-native hot path)")
    # NB type=bool keeps the reference's behaviour: any non-empty string is True
    p.add_argument("--seed_flag", type=bool, default=True, help="Fix random seed or not")
    p.add_argument("--seed", type=int, default=2024, help="random seed for init")
    p.add_argument("--cuda", type=bool, default=True, help="use gpu or not")
    p.add_argument("--gpu_id", type=int, default=0, help="gpu id")
    p.add_argument("--model", type=str, default="unknown", help="model name")
    return p.parse_args(argv)
