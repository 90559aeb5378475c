// stand-in: named by plan_env/map.h, nothing of it is used
