"""PPO on the GPU-resident AC environment: the reference's ``ac_solver/agents`` package
re-built around ``ACVectorEnv`` (rollout, reward wrappers and curriculum on the device), a fused
GAE / PPO-loss kernel pair (csrc/ppo_kernels.cu) and CUDA-graph captured rollout and minibatch steps."""
