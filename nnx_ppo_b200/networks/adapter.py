"""PPOAdapter — mirrors nnx_ppo/networks/adapter.py:61-133 (single action head, scalar value)."""
from __future__ import annotations

from typing import Any

from .types import ModuleState, StatefulModule


class PPOAdapter(StatefulModule):
    def __init__(self, action: StatefulModule, value: StatefulModule):
        self.action = action
        self.value = value

    def _children(self):
        return [self.action, self.value]

    def __call__(self, state, x: Any, rollout_extras: Any = None):
        from .plan import call_network
        return call_network(self, state, x, rollout_extras)

    def initialize_state(self, batch_size: int) -> dict[str, ModuleState]:
        return {"action": self.action.initialize_state(batch_size),
                "value": self.value.initialize_state(batch_size)}

    def reset_state(self, prev_state):
        return {"action": self.action.reset_state(prev_state["action"]),
                "value": self.value.reset_state(prev_state["value"])}

    def update_statistics(self, rollout_extras: Any) -> None:
        self.action.update_statistics(rollout_extras["action"])
        self.value.update_statistics(rollout_extras["value"])
