"""Entry point with the reference's command line (main.py:1-98):
    python main.py --model=LightGCN [--seed 2024 --gpu_id 0 --cuda True --seed_flag True]
reads ./configure/<Model>.txt, loads ./dataset/<name>/{train,test}.txt, logs to log/<Model>/<dataset>.log
and runs Trainer(args, config, dataset, device, logger).train().  The reference's own main.py also runs
unchanged from this directory (same importable module names); this file only drops the interactive menu
of models that are outside the accelerated hot path."""
import importlib
import logging
import os

import torch

import Parser
import utility.utility_data.data_loader as data_loader
import utility.utility_function.tools as tools

MODELS = ("LightGCN", "SimGCL", "XSimGCL", "NGCF", "MFBPR", "SGL", "LightCCF", "LightCSCF", "SCCF", "DirectAU", "EGCF")


def main(argv=None):
    args = Parser.parse_args(argv)
    if args.cuda:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(args.gpu_id)
    device = torch.device('cuda' if torch.cuda.is_available() else "cpu")
    if args.seed_flag:
        tools.set_seed(args.seed)
    if args.model not in MODELS:
        raise SystemExit("--model must be one of %s" % (MODELS,))
    Trainer = importlib.import_module("models." + args.model).Trainer
    config = tools.read_configuration('./configure/' + args.model + ".txt", args.model)
    os.makedirs('log/' + args.model, exist_ok=True)
    logger = logging.getLogger('logger')
    logger.setLevel(logging.INFO)
    logfile = logging.FileHandler('log/{}/{}.log'.format(args.model, config['dataset']), 'a', encoding='utf-8')
    logfile.setLevel(logging.INFO)
    logfile.setFormatter(logging.Formatter('%(asctime)s - %(message)s'))
    logger.addHandler(logfile)
    dataset = data_loader.Data(config['dataset_path'] + config['dataset'], config)
    logger.info("Run with " + args.model + " on " + config['dataset'])
    logger.info(dataset.get_statistics())
    recommender = Trainer(args, config, dataset, device, logger)
    for key in config:
        print("\t " + str(key) + " : " + str(config[key]))
        logger.info(str(key) + " : " + str(config[key]))
    recommender.train()


if __name__ == "__main__":
    main()
