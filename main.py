"""Unified entry point (drop-in for the reference's main.py): python main.py --algo mpnn --mode eval --steps 10 …"""
import argparse

from src.runner import Runner, RunnerArgs


def main(argv=None):
    parser = argparse.ArgumentParser(description="Unified runner for classical and RL experiments")
    parser.add_argument("--algo", choices=["dijkstra", "random", "mpnn", "mpnn+ppo"], default="random")
    parser.add_argument("--scenario", type=str, default="Easy")
    parser.add_argument("--mode", choices=["eval", "train"], default="eval")
    parser.add_argument("--timestep_size", type=int, default=1)
    parser.add_argument("--start-end-time", type=int, nargs=2, default=[0, 86400])
    parser.add_argument("--epochs", type=int, default=1)
    parser.add_argument("--rollout-steps", type=int, default=32)
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--device", type=str, default="cuda")
    parser.add_argument("--output-dir", type=str, default="runs")
    parser.add_argument("--profile", action="store_true")
    parser.add_argument("--torch-compile", action="store_true", help="accepted for compatibility; ignored")
    parser.add_argument("--steps", type=int, default=None, help="number of simulated timesteps in eval")
    parser.add_argument("--replicas", type=int, default=1, help="environment replicas per GPU for PPO rollouts")
    args = parser.parse_args(argv)
    runner = Runner(RunnerArgs(**vars(args)))
    runner.setup()
    if args.mode == "train":
        runner.train()
    runner.eval()
    return runner


if __name__ == "__main__":
    main()
