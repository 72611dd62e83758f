"""deepqnetwork_b200 -- B200-native DQN learner behind the reynkun/DeepQNetwork plugin surface.

The compute lives in ``libdqn_b200.so`` (hand-written sm_100a CUDA behind the C ABI declared in
``include/dqn_b200.h``); this package is the host-side mirror of the reference's Python interfaces
(``DeepQAgent``, ``DeepQModel`` and subclasses, ``ReplayMemory``, ``SumTree``, ``ReplaySampler``,
``ReplaySamplerPriority``).  There is no CPU fallback: importing the compute modules without the built
library, or creating a learner without a CUDA device, raises.
"""
__version__ = '0.1.0'
