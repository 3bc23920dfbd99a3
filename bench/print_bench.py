import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][0])
        print(path, "N=%d" % d["n_gpus"], "%.0f fps" % d["value"], "ms/step %.4f" % d["ms_per_step"], "wall %.4f" % d["wall_ms_per_step_incl_flush"],
              "e2e %.0f" % d["e2e"]["value"], "timeouts", d.get("sortfirst_wait_timeouts"))
    except Exception as e:
        print(path, "unreadable:", e)
