// stand-in: named by plan_env/src/risk_mapping_node.cpp, nothing of it is used
