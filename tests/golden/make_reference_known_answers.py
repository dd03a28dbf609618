"""Writes the known-answer vectors the reference's own unit tests hold for the hot path.

The numbers are the literals of the reference tests (cited per entry); nothing is computed by
our code.  Run:  python tests/golden/make_reference_known_answers.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

# CI/unit_tests/transformations/test_transformations.py:147-189 (test_unwrap_coordinates)
unwrap = {
    "source": "CI/unit_tests/transformations/test_transformations.py:147-189",
    "box": [1.1, 2.2, 3.3],
    # (A=1, T=4, 3) after the swapaxes at :161
    "pos": [[[0.5, 0.1, 3.2], [0.6, 2.1, 0.9], [0.6, 2.1, 2.1], [0.6, 0.1, 0.1]]],
    "last_pos": [[0.5, 0.1, 3.2]],
    "last_image_box": [[4, 0, 0]],
    "expected": [[[4 * 1.1 + 0.5, 0.1, 3.2], [4 * 1.1 + 0.6, -0.1, 4.2],
                  [4 * 1.1 + 0.6, -0.1, 5.4], [4 * 1.1 + 0.6, 0.1, 6.7]]],
    "expected_last_pos": [[0.6, 0.1, 0.1]],
    "expected_image_box": [[4, 0, 2]],
}
json.dump(unwrap, open(os.path.join(HERE, "unwrap_carry.json"), "w"), indent=1)

# CI/unit_tests/memory_manager/test_memory_manager.py:128-228 (planner arithmetic; the fake
# database of :28-46 returns (rows, columns, data_size); default scale = linear x10)
planner = {
    "source": "CI/unit_tests/memory_manager/test_memory_manager.py:128-228",
    "get_batch_size": [
        # (:141-148) and (:150-158)
        {"rows": 10, "columns": 10, "data_size": 500, "memory": 50000, "fraction": 0.5,
         "expect": [10, 1, 0]},
        {"rows": 11, "columns": 13, "data_size": 500, "memory": 50, "fraction": 1.0,
         "expect": [1, 13, 0]},
    ],
    # (:183-198): default fake database, data_range 5, machine memory (any realistic value)
    "atomwise_minibatch": {"rows": 10, "columns": 10, "data_size": 500, "data_range": 5,
                           "memory": 60e9, "fraction": 0.5,
                           "expect": {"batch_size": 10, "n_batches": 1, "n_atom_batches": 2,
                                      "atom_remainder": 0}},
    "get_ensemble_loop": [
        # batch_size, data_range, correlation_time -> (n_windows, minibatch)  (:200-228)
        {"batch_size": 50, "data_range": 10, "correlation_time": 5, "expect": [8, False]},
        {"batch_size": 5, "data_range": 10, "correlation_time": 5, "expect": [1, True]},
    ],
    # scale functions (:60-126)
    "scale_functions": [
        {"fn": {"quadratic": {"inner_scale_factor": 2, "outer_scale_factor": 2}}, "x": 10,
         "expect": 800},
        {"fn": {"polynomial": {"inner_scale_factor": 2, "outer_scale_factor": 2, "order": 3}},
         "x": 10, "expect": 16000},
    ],
}
json.dump(planner, open(os.path.join(HERE, "planner_cases.json"), "w"), indent=1)
print("written")
