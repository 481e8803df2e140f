"""td3_b200 -- B200-native (sm_100a) TD3 update hot path behind yannikkellerde/TD3's Python API.

Modules mirror the reference's names so its drivers import them unchanged:
``TD3_base``, ``TD3_featured``, ``TD3_particles``, ``my_replay_buffer``.
"""
__version__ = "0.1.0"
