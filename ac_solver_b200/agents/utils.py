"""Helpers of the PPO package (reference: ``ac_solver/agents/utils.py``)."""

from ..search.miller_schupp.miller_schupp import load_initial_states_from_text_file  # noqa: F401
