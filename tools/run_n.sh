# usage: bash tools/run_n.sh N script.py [args...]   — torchrun wrapper on one node
n=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) "$@"
