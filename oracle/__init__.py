"""CPU oracle for the FileBeep receive hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product (audio-modem-radio_b200/) never does:
it fails loudly when libfbdsp.so is missing.

Modules
  modem_v2   numpy/scipy restatement of the executable reference modem.py
             (pinned: tests/golden/*.npz generated from /root/reference itself)
  modem_v1   numpy restatement of the bytecode-only "v1" demodulators
             (SURVEY.md Appendix B).  PARITY UNPINNED: no executable reference.
  fec        restatement of fec.py decode (pinned by golden vectors)
  frames     restatement of decoder.parse_fbp_stream_enhanced + encoder._frame_data
  signals    seeded synthetic-signal generators (vectorised restatement of the
             reference modulators, validated against them)
"""
