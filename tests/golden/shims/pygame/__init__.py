"""Empty stand-in for `pygame` (only used inside MultiGridEnv.render, base.py:800-841)."""
