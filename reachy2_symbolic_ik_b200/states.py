"""State codes of the CUDA library <-> the reference's exact state strings."""
STATE_STRINGS = {
    0: "reachable",                          # symbolic_ik.py:234
    1: "Pose out of reach",                  # symbolic_ik.py:300
    2: "Backward pose",                      # symbolic_ik.py:306
    3: "wrist out of range",                 # symbolic_ik.py:159
    4: "limited by wrist",                   # symbolic_ik.py:262
    5: "out of reach - should not happen",   # symbolic_ik.py:281
    6: "limited by shoulder",                # control_ik.py:363,452
    7: "",                                   # control_ik.py:297 (continuous mode, success)
    8: "emergency",                          # text assembled by emergency_text()
    9: "invalid rotation",                   # scipy from_matrix ValueError (det <= 0)
}
STATE_CODES = {v: k for k, v in STATE_STRINGS.items()}

STATE_REACHABLE = 0
STATE_EMERGENCY = 8
STATE_INVALID_ROTATION = 9


def emergency_text(bits: int) -> str:
    """The emergency_state text the reference accumulates (utils.py:544-566)."""
    s = ""
    if bits & 1:
        s += "\n" + "EMERGENCY STOP: shoulder pitch limit reached"
    if bits & 2:
        s += "\n" + "EMERGENCY STOP: elbow yaw limit reached"
    if bits & 4:
        s += "\n" + "EMERGENCY STOP: wrist yaw limit reached"
    return s
