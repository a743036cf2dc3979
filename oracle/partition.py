"""TEST INFRASTRUCTURE ONLY: line-by-line Python restatement of the reference's `distribute_weights_swapping`
(MPMP.jl:425-465), 1-based cursors translated to 0-based. The product implementation is `clr::partition_weights`
(csrc/multi.cu, behind clrsdp_partition); tests/test_partition.py holds the two against each other."""


def distribute_weights_swapping(weights, n, nswaps=None):
    N = len(weights)
    nswaps = N * N if nswaps is None else nswaps            # (:425)
    step = N // n + 1                                        # (:428) first sets have this size
    nstep = n - (step * n - N)                               # (:429)
    sets = [list(range(i * step, (i + 1) * step)) for i in range(nstep)] + \
           [list(range(nstep * step + i * (step - 1), nstep * step + (i + 1) * (step - 1))) for i in range(n - nstep)]  # (:430-431)
    set_weights = [sum(weights[e] for e in s) for s in sets]  # (:432)
    index_set, index_el = 0, 0                               # (:433-434), 0-based
    for _ in range(nswaps):                                  # (:435)
        order = sorted(((set_weights[i], i) for i in range(n)), reverse=True)          # (:436)
        max_set = order[min(index_set, n - 1)][1]
        if not sets[max_set]:
            break
        eo = sorted(((weights[e], i) for i, e in enumerate(sets[max_set])), reverse=True)   # (:437)
        max_pos = eo[min(index_el, len(eo) - 1)][1]
        max_el = sets[max_set][max_pos]
        min_set = min(range(n), key=lambda i: set_weights[i])                          # argmin: first minimum (:438)
        if not sets[min_set]:
            break
        min_el = min(sets[min_set], key=lambda e: weights[e])                          # (:439) first minimum
        if (set_weights[min_set] + weights[max_el] - weights[min_el] < set_weights[max_set] and
                set_weights[max_set] - weights[max_el] + weights[min_el] < set_weights[max_set]):   # (:441-442)
            sets[max_set] = [e for e in sets[max_set] if e != max_el] + [min_el]       # (:444-445)
            set_weights[max_set] += weights[min_el] - weights[max_el]
            sets[min_set] = [e for e in sets[min_set] if e != min_el] + [max_el]       # (:448-449)
            set_weights[min_set] += weights[max_el] - weights[min_el]
            index_el, index_set = 0, 0                                                 # (:451-452)
        elif index_el + 1 < len(sets[min(index_set, n - 1)]):                          # (:453) NB: sets[index_set], as written
            index_el += 1
        elif index_el + 1 == step - 1 and index_set + 1 < n - 1:                       # (:455)
            index_set += 1
            index_el = 0
        else:
            break                                                                      # (:458-461)
    return sets, set_weights, [[weights[e] for e in s] for s in sets]                  # (:464)
