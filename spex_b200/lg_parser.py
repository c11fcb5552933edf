"""Command-line flags of the LightGCN_SPEX entry points: same names, types and defaults as
/root/reference/LightGCN_SPEX/code/lg_parser.py:3-24, plus --data_path being honoured by Loader
(the reference parses it but hard-codes "../data/", dataloader.py:74)."""
import argparse


def build_parser():
    p = argparse.ArgumentParser(description="Go lightGCN")
    p.add_argument("--cuda_id", default="0", help="which device to use")
    p.add_argument("--data_path", nargs="?", default="../data/", help="Input data path.")
    p.add_argument("--dataset", type=str, default="twitter", help="available datasets: [epinion2,weibo,twitter]")
    p.add_argument("--nb_heads", type=int, default=3, help="Number of head attentions.")
    p.add_argument("--recdim", type=int, default=64, help="the embedding size of lightGCN")
    p.add_argument("--layer", type=int, default=3, help="the layer num of lightGCN")
    p.add_argument("--lr", type=float, default=0.001, help="the learning rate")
    p.add_argument("--dropout", type=int, default=0, help="using the dropout or not")
    p.add_argument("--keepprob", type=float, default=0.6, help="edge keep probability of the dropout graph")
    p.add_argument("--a_fold", type=int, default=100, help="the fold num used to split large adj matrix")
    p.add_argument("--epochs", type=int, default=50)
    p.add_argument("--seed", type=int, default=2020, help="random seed")
    p.add_argument("--A_split", type=int, default=0, help="")
    p.add_argument("--batch_size", type=int, default=256, help="")
    p.add_argument("--batchSize", type=int, default=256, help="input batch size")
    p.add_argument("--hiddenSize", type=int, default=64, help="hidden state size")
    p.add_argument("--nonhybrid", action="store_true", help="only use the global preference to predict")
    p.add_argument("--act", type=int, default=1, help="activation function")
    return p


def parse_args_r(argv=None):
    return build_parser().parse_args(argv)
