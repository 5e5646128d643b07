// params.cpp -- parameter handling of the C ABI (include/mpc_b200.h): defaults, the flat
// mpc_params.yaml loader and the LoadParams string keys.  Host-only, no CUDA.
#include "../../include/mpc_b200.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

extern "C" {

void mpc_b200_params_default(mpc_b200_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    // MPC::MPC (mpc_ros/src/mpc_planner.cpp:227-230)
    p->mpc_steps = 20;
    p->max_angvel = 3.0;
    p->max_throttle = 1.0;
    p->bound_value = 1.0e3;
    // FG_eval::FG_eval (mpc_planner.cpp:47-57); its own _mpc_steps = 40 (:59) is always overwritten
    // by LoadParams(STEPS) before use, and MPC::Solve sizes the problem from MPC::_mpc_steps.
    p->dt = 0.1;
    p->ref_cte = 0.0; p->ref_etheta = 0.0; p->ref_vel = 0.5;
    p->w_cte = 100.0; p->w_etheta = 100.0; p->w_vel = 1.0;
    p->w_angvel = 100.0; p->w_accel = 50.0; p->w_angvel_d = 0.0; p->w_accel_d = 0.0;
    p->tol = 1e-8;
    p->max_iter = 100;
    // DrivingStateContext defaults (mpc_ros/src/driving_state.cpp:24-29) / MPCPlanner.cfg:14-20
    p->delay_mode = 1; p->max_speed = 0.7; p->path_length = 5.0; p->waypoints_dist = -1.0;
    p->goal_radius = 0.5; p->controller_freq = 10.0;
    p->warm_mu_init = 1e-3;
}

void mpc_b200_params_yaml_default(mpc_b200_params *p)
{
    if (!p) return;
    mpc_b200_params_default(p);
    // mpc_ros/params/mpc_params.yaml:2-25
    p->delay_mode = 1; p->max_speed = 0.5; p->waypoints_dist = -1.0; p->path_length = 5.0;
    p->goal_radius = 0.5; p->controller_freq = 10.0; p->dt = 1.0 / 10.0;
    p->mpc_steps = 20;
    p->ref_cte = 0.0; p->ref_vel = 0.5; p->ref_etheta = 0.0;
    p->w_cte = 100.0; p->w_etheta = 0.0; p->w_vel = 1000.0;
    p->w_angvel = 100.0; p->w_angvel_d = 0.0; p->w_accel = 50.0; p->w_accel_d = 0.0;
    p->max_angvel = 1.5; p->max_throttle = 1.0; p->bound_value = 1.0e3;
}

int mpc_b200_params_set(mpc_b200_params *p, const char *key, double v)
{
    if (!p || !key) return MPC_B200_ERR_INVALID;
    // keys of MPC::LoadParams / FG_eval::LoadParams (mpc_planner.cpp:73-85, :247-250)
    if (!strcmp(key, "DT")) p->dt = v;
    else if (!strcmp(key, "STEPS")) p->mpc_steps = (int32_t)v;     // int conversion as at :74, :247
    else if (!strcmp(key, "REF_CTE")) p->ref_cte = v;
    else if (!strcmp(key, "REF_ETHETA")) p->ref_etheta = v;
    else if (!strcmp(key, "REF_V")) p->ref_vel = v;
    else if (!strcmp(key, "W_CTE")) p->w_cte = v;
    else if (!strcmp(key, "W_EPSI")) p->w_etheta = v;
    else if (!strcmp(key, "W_V")) p->w_vel = v;
    else if (!strcmp(key, "W_ANGVEL")) p->w_angvel = v;
    else if (!strcmp(key, "W_A")) p->w_accel = v;
    else if (!strcmp(key, "W_DANGVEL")) p->w_angvel_d = v;
    else if (!strcmp(key, "W_DA")) p->w_accel_d = v;
    else if (!strcmp(key, "ANGVEL")) p->max_angvel = v;
    else if (!strcmp(key, "MAXTHR")) p->max_throttle = v;
    else if (!strcmp(key, "BOUND")) p->bound_value = v;
    else return MPC_B200_ERR_INVALID;
    return MPC_B200_OK;
}

static int parse_bool(const char *s, int *out)
{
    if (!strncmp(s, "true", 4) || !strncmp(s, "True", 4) || !strncmp(s, "TRUE", 4)) { *out = 1; return 1; }
    if (!strncmp(s, "false", 5) || !strncmp(s, "False", 5) || !strncmp(s, "FALSE", 5)) { *out = 0; return 1; }
    return 0;
}

int mpc_b200_params_from_yaml(const char *path, mpc_b200_params *p)
{
    if (!path || !p) return MPC_B200_ERR_INVALID;
    FILE *f = fopen(path, "r");
    if (!f) return MPC_B200_ERR_IO;
    char line[512];
    while (fgets(line, sizeof(line), f)) {
        char *hash = strchr(line, '#');
        if (hash) *hash = 0;
        char *colon = strchr(line, ':');
        if (!colon) continue;
        *colon = 0;
        char *k = line;
        while (*k && isspace((unsigned char)*k)) k++;
        char *ke = k + strlen(k);
        while (ke > k && isspace((unsigned char)ke[-1])) *--ke = 0;
        char *v = colon + 1;
        while (*v && isspace((unsigned char)*v)) v++;
        if (!*k || !*v) continue;
        int b;
        if (!strcmp(k, "delay_mode")) { if (parse_bool(v, &b)) p->delay_mode = b; continue; }
        if (parse_bool(v, &b)) continue;   // pub_twist_cmd, debug_info: not solver parameters
        char *end = NULL;
        const double x = strtod(v, &end);
        if (end == v) continue;
        if (!strcmp(k, "mpc_steps")) p->mpc_steps = (int32_t)x;
        else if (!strcmp(k, "mpc_ref_cte")) p->ref_cte = x;
        else if (!strcmp(k, "mpc_ref_vel")) p->ref_vel = x;
        else if (!strcmp(k, "mpc_ref_etheta")) p->ref_etheta = x;
        else if (!strcmp(k, "mpc_w_cte")) p->w_cte = x;
        else if (!strcmp(k, "mpc_w_etheta")) p->w_etheta = x;
        else if (!strcmp(k, "mpc_w_vel")) p->w_vel = x;
        else if (!strcmp(k, "mpc_w_angvel")) p->w_angvel = x;
        else if (!strcmp(k, "mpc_w_angvel_d")) p->w_angvel_d = x;
        else if (!strcmp(k, "mpc_w_accel")) p->w_accel = x;
        else if (!strcmp(k, "mpc_w_accel_d")) p->w_accel_d = x;
        else if (!strcmp(k, "mpc_max_angvel")) p->max_angvel = x;
        else if (!strcmp(k, "mpc_max_throttle")) p->max_throttle = x < 0.1 ? 0.1 : x;   // floor of driving_state.cpp:63
        else if (!strcmp(k, "mpc_bound_value")) p->bound_value = x;
        else if (!strcmp(k, "controller_freq")) { p->controller_freq = x; if (x > 0.0) p->dt = 1.0 / x; }
        else if (!strcmp(k, "max_speed")) p->max_speed = x;
        else if (!strcmp(k, "path_length")) p->path_length = x;
        else if (!strcmp(k, "waypoints_dist")) p->waypoints_dist = x;
        else if (!strcmp(k, "goal_radius")) p->goal_radius = x;
    }
    fclose(f);
    return MPC_B200_OK;
}

}  // extern "C"
