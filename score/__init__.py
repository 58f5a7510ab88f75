"""Drop-in alias package: ``from score.solve_score import solve_score`` resolves to the
B200-native implementation in ``score_b200`` (same module layout as the reference)."""
