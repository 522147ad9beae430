import sys, json
sys.path.insert(0, "/root/repo")
import optimal_control_problem_b200 as ocp
prob = ocp.Problem("quadrotor")
print(json.dumps(prob.solver.launch_plan()))
