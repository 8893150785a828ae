/*
 * oracle/ppe_oracle.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Plain-C, sequential restatement of the hot path of afb2001/path_planner, one edge at a time,
 * exactly as the reference computes it.  Each function cites the reference lines it follows.
 * The Dubins solver/sampler it calls is oracle/dubins.c (restatement of the un-vendored
 * dubins_curves dependency).
 *
 * Pinning: this restatement is checked (tests/test_oracle_vs_ref.py, run where /root/reference
 * exists) against oracle/_ref/libref_planner.so = the reference's own Edge.cpp / Vertex.cpp /
 * Ribbon*.cpp / *DynamicObstaclesManager.cpp / GridWorldMap.cpp compiled from /root/reference,
 * against the reference's known-answer tests transcribed in tests/test_golden_reference.py, and
 * against committed fixtures in tests/golden/ generated from that compiled reference.
 * What stays unpinned is the dubins_curves arithmetic itself (see dubins.h).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use this file.
 */
#include <math.h>
#include <float.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "dubins.h"
#include "ppe.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif

typedef struct { double sx, sy, ex, ey; } rib_t;

typedef struct {
    int n;
    rib_t* r;
    double cct; /* RibbonManager::m_CoverageCompletedTime */
} ribset_t;

typedef struct oracle_ctx {
    ppe_config cfg;
    int have_cfg;
    /* Map / GridWorldMap */
    int map_kind; /* 0 none, 1 bitmap */
    uint8_t* bits;
    int rows, cols, stride;
    double res;
    /* obstacle managers: 9 doubles each: X Y Yaw Speed Time + (Width Length 0 0 | cov00 cov01 cov10 cov11) */
    int obs_kind; /* 0 base, 1 binary, 2 gaussian */
    int n_obs;
    double* obs;
    /* interned ribbon sets */
    ribset_t* sets;
    int n_sets, cap_sets;
    /* ribbons-after of the last batch */
    rib_t** after;
    int* n_after;
    int64_t n_last;
    char err[256];
} oracle_ctx;

#define RIB_CAP 4096

/* ------------------------------------------------------------------------------------------- */
/* Ribbon (path_planner/src/planner/utilities/Ribbon.{h,cpp})                                    */
/* ------------------------------------------------------------------------------------------- */
static const double RIB_TOL = 1e-5;      /* Ribbon.h:129 c_Tolerance */
static const double RIB_STRICT = 2;      /* Ribbon.h:131 c_StrictModifier */

/* Ribbon.h:134-136 */
static double rib_sqlen(const rib_t* r) {
    return (r->ex - r->sx) * (r->ex - r->sx) + (r->ey - r->sy) * (r->ey - r->sy);
}
/* Ribbon.cpp:52-58 */
static double rib_min_length(double W) { return 2 * W; }
/* Ribbon.cpp:23-25 */
static int rib_covered(const rib_t* r, int strict, double W) {
    return rib_sqlen(r) < rib_min_length(W) * rib_min_length(W) / (strict ? RIB_STRICT * RIB_STRICT : 1);
}
/* Ribbon.cpp:27-29 */
static double rib_length(const rib_t* r) { return sqrt(rib_sqlen(r)); }
/* Ribbon.cpp:72-78 */
static void rib_projection(const rib_t* r, double x, double y, double* px, double* py) {
    double squaredL = rib_sqlen(r);
    double dot = (x - r->sx) * (r->ex - r->sx) + (y - r->sy) * (r->ey - r->sy);
    double projectedX = (r->ex - r->sx) * dot / squaredL;
    double projectedY = (r->ey - r->sy) * dot / squaredL;
    *px = projectedX + r->sx;
    *py = projectedY + r->sy;
}
/* Ribbon.cpp:90-95 */
static int rib_contains_projection(const rib_t* r, double px, double py) {
    return !(((px - r->sx < -RIB_TOL && px - r->ex < -RIB_TOL) || (px - r->sx > RIB_TOL && px - r->ex > RIB_TOL)) ||
             ((py - r->sy < -RIB_TOL && py - r->ey < -RIB_TOL) || (py - r->sy > RIB_TOL && py - r->ey > RIB_TOL)));
}
/* Ribbon.h:118-121 */
static double rib_distance(const rib_t* r, double x, double y) {
    return (fabs((r->ey - r->sy) * x - (r->ex - r->sx) * y + r->ex * r->sy - r->ey * r->sx)) / sqrt(rib_sqlen(r));
}
/* Ribbon.cpp:39-43 */
static int rib_contains(const rib_t* r, double x, double y, double px, double py, int strict, double W) {
    double d;
    if (!rib_contains_projection(r, px, py)) return 0;
    d = rib_distance(r, x, y);
    return d < (strict ? W / RIB_STRICT : W);
}
/* Ribbon.cpp:9-17; returns the split-off piece (empty ribbon 0,0,0,0 when not contained) */
static rib_t rib_split(rib_t* r, double x, double y, int strict, double W) {
    double px, py;
    rib_t piece = {0, 0, 0, 0};
    rib_projection(r, x, y, &px, &py);
    if (!rib_contains(r, x, y, px, py, strict, W)) return piece;
    piece.sx = r->sx; piece.sy = r->sy; piece.ex = px; piece.ey = py;
    r->sx = px; r->sy = py;
    return piece;
}

/* ------------------------------------------------------------------------------------------- */
/* RibbonManager (path_planner/src/planner/utilities/RibbonManager.cpp)                          */
/* ------------------------------------------------------------------------------------------- */
static double pt_distance(double x1, double y1, double x2, double y2) { /* RibbonManager.h:289-291 */
    return sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
}

/* RibbonManager.cpp:14-22 (+ add, :154-158).  Returns -1 on capacity overflow. */
static int ribs_cover(rib_t* ribs, int* n, int cap, double x, double y, int strict, double W) {
    int i = 0;
    while (i < *n) {
        rib_t piece = rib_split(&ribs[i], x, y, strict, W);
        if (!rib_covered(&piece, strict, W)) {
            /* insert before i */
            if (*n >= cap) return -1;
            memmove(&ribs[i + 1], &ribs[i], (size_t)(*n - i) * sizeof(rib_t));
            ribs[i] = piece;
            (*n)++;
            i++; /* i keeps pointing at the (shortened) original ribbon */
        }
        if (rib_covered(&ribs[i], strict, W)) {
            memmove(&ribs[i], &ribs[i + 1], (size_t)(*n - i - 1) * sizeof(rib_t));
            (*n)--;
        } else {
            i++;
        }
    }
    return 0;
}

/* RibbonManager.cpp:142-152 */
static double ribs_min_distance_from(const rib_t* ribs, int n, double x, double y, double W) {
    double min = DBL_MAX;
    int i;
    if (n == 0) return 0;
    for (i = 0; i < n; i++) {
        double px, py, dStart, dEnd;
        rib_projection(&ribs[i], x, y, &px, &py);
        if (rib_contains(&ribs[i], x, y, px, py, 0, W)) return 0;
        dStart = pt_distance(ribs[i].sx, ribs[i].sy, x, y);
        dEnd = pt_distance(ribs[i].ex, ribs[i].ey, x, y);
        min = fmin(fmin(min, dEnd), dStart);
    }
    return min;
}

/* RibbonManager.cpp:234-248 */
static double ribs_max_distance(const rib_t* ribs, int n, double x, double y, double W) {
    double sumLength = 0, min = DBL_MAX, max = 0;
    int i;
    for (i = 0; i < n; i++) {
        double dStart, dEnd;
        sumLength += rib_length(&ribs[i]) - 2 * W;
        dStart = pt_distance(ribs[i].sx, ribs[i].sy, x, y);
        dEnd = pt_distance(ribs[i].ex, ribs[i].ey, x, y);
        min = fmin(fmin(min, dEnd), dStart);
        max = fmax(fmax(max, dEnd), dStart);
    }
    return fmax(sumLength + min, max);
}

/* ------------------------------------------------------------------------------------------- */
/* Map::isBlocked (Map.cpp:4-6) / GridWorldMap::isBlocked (GridWorldMap.cpp:84-93)               */
/* ------------------------------------------------------------------------------------------- */
/* RibbonManager::tspPointRobotNoSplitAllRibbons (RibbonManager.cpp:53-67) and ...KRibbons (:69-95): depth-first over
 * the ribbons left, both directions of each; the K variant first sorts the list it was handed with std::list::sort (stable)
 * by "nearest end point farther than" (the comparator's own comment says it should be the other way round) and tries
 * the first K.  `order` holds the ribbons left in list order. */
static double ribs_tsp_point_robot(const rib_t* ribs, const int* order, int n, double soFar, double px, double py, int K, double W) {
    int ord[64];
    double key[64];
    double mn = DBL_MAX;
    int i, j, tried = 0;
    if (n == 0) return soFar;
    for (i = 0; i < n; i++) ord[i] = order[i];
    if (K > 0) { /* stable insertion sort, comp(a, b) = key(a) > key(b) */
        for (i = 0; i < n; i++)
            key[i] = fmin(pt_distance(px, py, ribs[ord[i]].sx, ribs[ord[i]].sy), pt_distance(px, py, ribs[ord[i]].ex, ribs[ord[i]].ey));
        for (i = 1; i < n; i++) {
            const int o = ord[i];
            const double kv = key[i];
            j = i;
            while (j > 0 && kv > key[j - 1]) { ord[j] = ord[j - 1]; key[j] = key[j - 1]; j--; }
            ord[j] = o; key[j] = kv;
        }
    }
    for (i = 0; i < n; i++) {
        int rest[64];
        const rib_t* r;
        int m = 0;
        if (K > 0 && tried++ >= K) break;
        r = &ribs[ord[i]];
        for (j = 0; j < n; j++) if (j != i) rest[m++] = ord[j];
        mn = fmin(mn, ribs_tsp_point_robot(ribs, rest, m, fmax(soFar + rib_length(r) - 2 * W + pt_distance(px, py, r->sx, r->sy), 0),
                                           r->ex, r->ey, K, W));
        mn = fmin(mn, ribs_tsp_point_robot(ribs, rest, m, fmax(soFar + rib_length(r) - 2 * W + pt_distance(px, py, r->ex, r->ey), 0),
                                           r->sx, r->sy, K, W));
    }
    return mn;
}

static int map_blocked(const oracle_ctx* c, double x, double y) {
    size_t r, col;
    if (c->map_kind == 0) return 0;
    if (x < 0 || x / c->res >= (double)(size_t)c->cols) return 1;
    if (y < 0 || y / c->res >= (double)(size_t)c->rows) return 1;
    r = (size_t)(y / c->res);
    col = (size_t)(x / c->res);
    return (c->bits[r * (size_t)c->stride + (col >> 3)] >> (col & 7)) & 1;
}

/* ------------------------------------------------------------------------------------------- */
/* DynamicObstaclesManager::collisionExists                                                      */
/* ------------------------------------------------------------------------------------------- */
static double collision_exists(const oracle_ctx* c, double x, double y, double time, int strict) {
    double sum = 0;
    int i;
    if (c->obs_kind == 0) return 0; /* DynamicObstaclesManager.h:23 */
    if (c->obs_kind == 1) {
        /* BinaryDynamicObstaclesManager.cpp:4-22, Obstacle::project Binary...h:19-24 */
        for (i = 0; i < c->n_obs; i++) {
            const double* o = c->obs + 9 * i;
            double X = o[0], Y = o[1], Yaw = o[2], Speed = o[3], Time = o[4], Width = o[5], Length = o[6];
            double dt, dx, dy, translatedX, translatedY, rotatedX, rotatedY;
            if (strict) { Width += 2; Length += 2; }
            dt = time - Time;
            dx = Speed * dt * cos(Yaw);
            dy = Speed * dt * sin(Yaw);
            X += dx; Y += dy;
            translatedX = x - X;
            translatedY = y - Y;
            rotatedX = translatedX * cos(Yaw) - translatedY * sin(Yaw);
            rotatedY = translatedX * sin(Yaw) + translatedY * cos(Yaw);
            if (fabs(rotatedX) < Length / 2 && fabs(rotatedY) < Width / 2) sum++;
        }
        return sum;
    }
    /* GaussianDynamicObstaclesManager.cpp:3-13, Obstacle::project/pdf Gaussian...h:31-44;
     * 2x2 algebra as defined by oracle/include/eigen3/Eigen/Core */
    for (i = 0; i < c->n_obs; i++) {
        const double* o = c->obs + 9 * i;
        double X = o[0], Y = o[1], Yaw = o[2], Speed = o[3], Time = o[4];
        double c00 = o[5], c01 = o[6], c10 = o[7], c11 = o[8];
        double dt = time - Time;
        double dx = Speed * dt * cos(Yaw);
        double dy = Speed * dt * sin(Yaw);
        double twoPi = 2 * M_PI;
        double d0, d1, det, invdet, i00, i10, i01, i11, r0, r1, quadform, norm;
        X += dx; Y += dy;
        d0 = x - X; d1 = y - Y;
        det = c00 * c11 - c10 * c01;
        invdet = 1.0 / det;
        i00 = c11 * invdet; i10 = -c10 * invdet; i01 = -c01 * invdet; i11 = c00 * invdet;
        r0 = d0 * i00 + d1 * i10;
        r1 = d0 * i01 + d1 * i11;
        quadform = r0 * d0 + r1 * d1;
        norm = 1.0 / twoPi / sqrt(det);
        sum += norm * exp(-0.5 * quadform);
    }
    if (sum < 1e-5) return 0;
    return sum;
}

/* ------------------------------------------------------------------------------------------- */
/* State / DubinsWrapper (path_planner_common)                                                   */
/* ------------------------------------------------------------------------------------------- */
/* State::yaw, State.h:51-55 */
static double state_yaw(double heading) {
    double h = M_PI_2 - heading;
    if (h < 0) h += 2 * M_PI;
    return h;
}

typedef struct {
    DubinsPath path;
    double speed, start, end, ustart; /* m_Speed, m_StartTime, m_EndTime, m_UpdatedStartTime */
} wrapper_t;

static void wrapper_init(wrapper_t* w) { /* DubinsWrapper.h:118-120 default members */
    memset(w, 0, sizeof *w);
    w->start = -1; w->end = -1; w->ustart = -1;
}

/* DubinsWrapper::sample, DubinsWrapper.cpp:29-49.  pose = {x, y, heading, speed}.
 * Returns 1 where the reference throws std::runtime_error. */
static int wrapper_sample(const wrapper_t* w, double pose[4], double time) {
    double distance;
    int err;
    if (!(w->start >= 0)) return 1;                      /* containsTime on an unset wrapper, :24-26 */
    if (!(w->ustart <= time && w->end >= time)) return 1; /* :30-35 */
    distance = (time - w->start) * w->speed;
    err = dubins_path_sample(&w->path, distance, pose);
    if (err == EDUBPARAM) err = dubins_path_sample(&w->path, distance - 1e-5, pose);
    /* State::setYaw(s.heading()), State.h:62-65 */
    pose[2] = M_PI_2 - pose[2];
    if (pose[2] < 0) pose[2] += 2 * M_PI;
    pose[3] = w->speed;
    return 0;
}

/* ------------------------------------------------------------------------------------------- */
/* Edge::computeTrueCost, Edge.cpp:68-206 (+ computeApproxCost :11-20, setEnd :208-215,          */
/* Vertex::setCurrentCost Vertex.cpp:102-104, Vertex::computeApproxToGo Vertex.cpp:49-64)        */
/* ------------------------------------------------------------------------------------------- */
static void true_cost_one(const oracle_ctx* c, const ppe_edge* e, ppe_edge_result* r, rib_t* ribs, int* n_ribs) {
    const ppe_config* cfg = &c->cfg;
    const double W = cfg->ribbon_width;
    const double inc = cfg->collision_checking_increment;
    const ribset_t* parent = &c->sets[e->ribbon_set];
    wrapper_t w;
    double endState[4]; /* end()->state() pose */
    double approx = -1;
    double speed, rho;
    double cct = parent->cct;
    int n = parent->n;
    int infeasible = 0;
    int i;

    memset(r, 0, sizeof *r);
    r->ribbons_offset = -1;
    memcpy(ribs, parent->r, (size_t)n * sizeof(rib_t)); /* v->m_RibbonManager = start->m_RibbonManager, Vertex.cpp:24,32 */
    wrapper_init(&w);

    if (e->has_path) {
        /* Vertex::connect(start, wrapper, coverageAllowed) -> Edge::setEnd(wrapper), Edge.cpp:208-215 */
        w.path.qi[0] = e->path_qi[0]; w.path.qi[1] = e->path_qi[1]; w.path.qi[2] = e->path_qi[2];
        w.path.param[0] = e->path_param[0]; w.path.param[1] = e->path_param[1]; w.path.param[2] = e->path_param[2];
        w.path.rho = e->path_rho; w.path.type = (DubinsPathType)e->path_type;
        w.speed = e->w_speed; w.start = w.ustart = e->w_start_time;
        w.end = w.start + dubins_path_length(&w.path) / w.speed;      /* fill(), DubinsWrapper.cpp:84-93 */
        if (e->w_end_time < w.end) w.end = e->w_end_time;             /* earlier updateEndTime, :100-104 */
        endState[0] = endState[1] = endState[2] = endState[3] = 0;
        if (wrapper_sample(&w, endState, w.end)) { r->status = PPE_EDGE_ERR_END_SAMPLE; return; }
        approx = (w.end - e->src[4]) * cfg->time_penalty_factor;
    } else {
        endState[0] = e->dst[0]; endState[1] = e->dst[1]; endState[2] = e->dst[2]; endState[3] = e->dst[3];
    }

    /* Edge.cpp:73-84 */
    speed = endState[3];
    rho = e->coverage_allowed ? cfg->coverage_turning_radius : cfg->turning_radius;
    if (approx == -1 || w.path.rho != rho) {
        /* computeApproxCost(speed, rho), Edge.cpp:11-20 */
        if (e->src[0] == endState[0] && e->src[1] == endState[1] && e->src[2] == endState[2]) {
            approx = 0; /* co-located: wrapper left as it was */
        } else {
            /* DubinsWrapper::set, DubinsWrapper.cpp:9-17 */
            double q1[3], q2[3];
            q1[0] = e->src[0]; q1[1] = e->src[1]; q1[2] = state_yaw(e->src[2]);
            q2[0] = endState[0]; q2[1] = endState[1]; q2[2] = state_yaw(endState[2]);
            dubins_shortest_path(&w.path, q1, q2, rho);
            w.speed = e->src[3];
            w.ustart = w.start = e->src[4];
            w.end = w.start + dubins_path_length(&w.path) / w.speed;
            approx = dubins_path_length(&w.path) / speed * cfg->time_penalty_factor;
        }
    }
    if (w.speed != speed) {
        /* setSpeed -> setEndTime -> length() throws on an unset wrapper (DubinsWrapper.cpp:19-22,120-123) */
        if (!(w.start >= 0)) { r->status = PPE_EDGE_ERR_NO_PATH; return; }
        w.speed = speed;
        w.end = w.start + dubins_path_length(&w.path) / w.speed;
    }
    if (approx < 0) { r->status = PPE_EDGE_ERR_NO_PATH; return; } /* Edge.cpp:85 */

    {
        double collisionPenalty = 0;
        double P[4]; /* `intermediate` */
        double t = e->src[4];
        double endTime = fmin(cfg->time_horizon + 1e-12 + cfg->start_state_time, w.end); /* :90 */
        int ribbonsDoneTime = -1;                       /* `auto x = -1` is an int, :92 */
        int startedDone = (n == 0);                     /* :93 */
        double toCoverDistance = 0;
        double lastHeading = e->src[2];
        double timeIncrement, timeSinceStart, timeNudge, T, netTime;
        int n_samples = 0, n_checkpoints = 0;

        P[0] = e->src[0]; P[1] = e->src[1]; P[2] = e->src[2]; P[3] = e->src[3];
        if (t >= endTime) infeasible = 1;               /* :102-110 */
        timeIncrement = inc / cfg->max_speed;           /* :114 */
        timeSinceStart = t - cfg->start_state_time;     /* :118-120 */
        timeNudge = fmod(timeSinceStart, timeIncrement);
        t += timeNudge;

        while (t < endTime) {                           /* :125 */
            n_samples++;
            if (wrapper_sample(&w, P, t)) { infeasible = 1; break; }     /* :126-133 */
            if (map_blocked(c, P[0], P[1])) { infeasible = 1; break; }  /* :144-147 */
            collisionPenalty += collision_exists(c, P[0], P[1], t, 1) * cfg->collision_penalty_factor; /* :150-151 */
            if (toCoverDistance > inc) {                /* :153-154 */
                toCoverDistance -= inc;
            } else {
                n_checkpoints++;
                toCoverDistance = ribs_min_distance_from(ribs, n, P[0], P[1], W);        /* :158 */
                if (e->coverage_allowed || lastHeading == P[2]) {                        /* :159-161 */
                    if (ribs_cover(ribs, &n, RIB_CAP, P[0], P[1], 1, W)) { r->status = PPE_EDGE_ERR_RIBBON_CAPACITY; return; }
                }
                if (n == 0) {                           /* :162-170 */
                    if (cct == -1) cct = t;
                    ribbonsDoneTime = (int)t;
                    endTime = fmin(endTime, cct + cfg->time_minimum);
                }
            }
            t += timeIncrement;                         /* :173-174 */
            lastHeading = P[2];
        }
        /* :177-179 */
        if (wrapper_sample(&w, endState, endTime)) {
            r->status = PPE_EDGE_ERR_END_SAMPLE;
            r->infeasible = infeasible;
            return;
        }
        w.end = endTime;
        /* :182-191 */
        if (e->coverage_allowed || lastHeading == P[2]) {
            if (ribs_cover(ribs, &n, RIB_CAP, P[0], P[1], 1, W)) { r->status = PPE_EDGE_ERR_RIBBON_CAPACITY; return; }
        }
        if (n == 0) {
            if (cct == -1) cct = t;
            ribbonsDoneTime = (int)t;
        }
        /* :195-199 */
        netTime = endTime - e->src[4];
        T = fmax(netTime - (n == 0 ? (endTime - ribbonsDoneTime) : 0), 0);
        if (startedDone) T = 0;
        r->collision_penalty = collisionPenalty;
        r->true_cost = T * cfg->time_penalty_factor + collisionPenalty;
        r->n_samples = n_samples;
        r->n_checkpoints = n_checkpoints;
        r->end[4] = endTime;
    }

    r->approx_cost = approx;
    r->infeasible = infeasible;
    r->end[0] = endState[0]; r->end[1] = endState[1]; r->end[2] = endState[2]; r->end[3] = endState[3];
    r->g = e->src_g + r->true_cost;                                              /* Vertex.cpp:102-104 */
    if (cfg->heuristic == PPE_H_MAX_DISTANCE) {                                  /* Vertex.cpp:49-64, RibbonManager.cpp:28-51 */
        double d = n == 0 ? 0 : ribs_max_distance(ribs, n, r->end[0], r->end[1], W);
        r->h = d / cfg->max_speed * cfg->time_penalty_factor;
    } else if ((cfg->heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_ALL || cfg->heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K) && n <= 8) {
        int order[8];
        double d;
        for (i = 0; i < n; i++) order[i] = i;
        d = n == 0 ? 0 : ribs_tsp_point_robot(ribs, order, n, 0, r->end[0], r->end[1],
                                              cfg->heuristic == PPE_H_TSP_POINT_ROBOT_NO_SPLIT_K ? (cfg->tsp_k > 0 ? cfg->tsp_k : 2) : 0, W);
        r->h = d / cfg->max_speed * cfg->time_penalty_factor;
    } else {
        r->h = -1; /* Dubins TSP variants, lists longer than the device serves */
    }
    r->coverage_completed_time = cct;
    r->path_qi[0] = w.path.qi[0]; r->path_qi[1] = w.path.qi[1]; r->path_qi[2] = w.path.qi[2];
    r->path_param[0] = w.path.param[0]; r->path_param[1] = w.path.param[1]; r->path_param[2] = w.path.param[2];
    r->path_rho = w.path.rho; r->path_type = (int32_t)w.path.type;
    r->w_speed = w.speed; r->w_start_time = w.ustart; r->w_end_time = w.end;
    r->n_ribbons_after = n;
    r->ribbons_changed = (n != parent->n);
    if (!r->ribbons_changed)
        for (i = 0; i < n; i++)
            if (memcmp(&ribs[i], &parent->r[i], sizeof(rib_t)) != 0) { r->ribbons_changed = 1; break; }
    *n_ribs = n;
}

/* ------------------------------------------------------------------------------------------- */
/* C ABI (same shapes as include/ppe.h, prefix oracle_)                                          */
/* ------------------------------------------------------------------------------------------- */
int oracle_create(oracle_ctx** out) {
    oracle_ctx* c = (oracle_ctx*)calloc(1, sizeof(oracle_ctx));
    if (!c) return PPE_ERR_INVALID;
    *out = c;
    return PPE_OK;
}

static void free_after(oracle_ctx* c) {
    int64_t i;
    if (c->after) {
        for (i = 0; i < c->n_last; i++) free(c->after[i]);
        free(c->after); free(c->n_after);
    }
    c->after = NULL; c->n_after = NULL; c->n_last = 0;
}

int oracle_clear_ribbon_sets(oracle_ctx* c) {
    int i;
    for (i = 0; i < c->n_sets; i++) free(c->sets[i].r);
    c->n_sets = 0;
    return PPE_OK;
}

void oracle_destroy(oracle_ctx* c) {
    if (!c) return;
    oracle_clear_ribbon_sets(c);
    free(c->sets); free(c->bits); free(c->obs);
    free_after(c);
    free(c);
}

const char* oracle_last_error(const oracle_ctx* c) { return c->err; }

int oracle_set_config(oracle_ctx* c, const ppe_config* cfg) { c->cfg = *cfg; c->have_cfg = 1; return PPE_OK; }
const ppe_config* oracle_config(const oracle_ctx* c) { return &c->cfg; }

int oracle_set_map_none(oracle_ctx* c) { c->map_kind = 0; return PPE_OK; }

int oracle_set_map_bitmap(oracle_ctx* c, const uint8_t* bits, int rows, int cols, int stride, double res) {
    free(c->bits);
    c->bits = (uint8_t*)malloc((size_t)rows * stride);
    memcpy(c->bits, bits, (size_t)rows * stride);
    c->rows = rows; c->cols = cols; c->stride = stride; c->res = res; c->map_kind = 1;
    return PPE_OK;
}

int oracle_set_obstacles_none(oracle_ctx* c) { c->obs_kind = 0; c->n_obs = 0; return PPE_OK; }

int oracle_set_obstacles_binary(oracle_ctx* c, int n, const double* x, const double* y, const double* yaw,
                                const double* speed, const double* time, const double* width, const double* length) {
    int i;
    free(c->obs);
    c->obs = (double*)calloc((size_t)(n > 0 ? n : 1) * 9, sizeof(double));
    for (i = 0; i < n; i++) {
        double* o = c->obs + 9 * i;
        o[0] = x[i]; o[1] = y[i]; o[2] = yaw[i]; o[3] = speed[i]; o[4] = time[i]; o[5] = width[i]; o[6] = length[i];
    }
    c->n_obs = n; c->obs_kind = 1;
    return PPE_OK;
}

int oracle_set_obstacles_gaussian(oracle_ctx* c, int n, const double* x, const double* y, const double* yaw,
                                  const double* speed, const double* time, const double* cov) {
    int i;
    free(c->obs);
    c->obs = (double*)calloc((size_t)(n > 0 ? n : 1) * 9, sizeof(double));
    for (i = 0; i < n; i++) {
        double* o = c->obs + 9 * i;
        o[0] = x[i]; o[1] = y[i]; o[2] = yaw[i]; o[3] = speed[i]; o[4] = time[i];
        if (cov) { o[5] = cov[4 * i]; o[6] = cov[4 * i + 1]; o[7] = cov[4 * i + 2]; o[8] = cov[4 * i + 3]; }
        else { o[5] = 30; o[6] = 10; o[7] = 10; o[8] = 30; } /* Gaussian...h:24-25 */
    }
    c->n_obs = n; c->obs_kind = 2;
    return PPE_OK;
}

int oracle_put_ribbon_set(oracle_ctx* c, int n, const double* xyxy, double cct, int32_t* id) {
    ribset_t* s;
    int i;
    if (!c->have_cfg) { snprintf(c->err, sizeof c->err, "set_config first (ribbon width)"); return PPE_ERR_STATE; }
    if (c->n_sets == c->cap_sets) {
        c->cap_sets = c->cap_sets ? 2 * c->cap_sets : 16;
        c->sets = (ribset_t*)realloc(c->sets, (size_t)c->cap_sets * sizeof(ribset_t));
    }
    s = &c->sets[c->n_sets];
    s->r = (rib_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(rib_t));
    s->n = 0;
    for (i = 0; i < n; i++) {
        rib_t rb;
        rb.sx = xyxy[4 * i]; rb.sy = xyxy[4 * i + 1]; rb.ex = xyxy[4 * i + 2]; rb.ey = xyxy[4 * i + 3];
        /* verbatim: a child vertex copies its parent's list (Vertex.cpp:24,32); the covered-filter of
         * RibbonManager::add (RibbonManager.cpp:154-158) has been applied by whoever built the list */
        s->r[s->n++] = rb;
    }
    s->cct = cct;
    *id = c->n_sets++;
    return PPE_OK;
}

int oracle_dubins_batch(oracle_ctx* c, int64_t n, const double* q0, const double* q1, const double* rho,
                        int32_t* type, double* param, double* length, int32_t* err) {
    int64_t i;
    (void)c;
    for (i = 0; i < n; i++) {
        DubinsPath p;
        double a[3], b[3];
        memset(&p, 0, sizeof p);
        a[0] = q0[3 * i]; a[1] = q0[3 * i + 1]; a[2] = q0[3 * i + 2];
        b[0] = q1[3 * i]; b[1] = q1[3 * i + 1]; b[2] = q1[3 * i + 2];
        err[i] = dubins_shortest_path(&p, a, b, rho[i]);
        type[i] = (int32_t)p.type;
        param[3 * i] = p.param[0]; param[3 * i + 1] = p.param[1]; param[3 * i + 2] = p.param[2];
        length[i] = err[i] == EDUBOK ? dubins_path_length(&p) : 0;
    }
    return PPE_OK;
}

typedef struct {
    oracle_ctx* c;
    const ppe_edge* edges;
    ppe_edge_result* results;
    int64_t n;
    int keep;
    int64_t next; /* guarded by mu */
    pthread_mutex_t mu;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    rib_t* ribs = (rib_t*)malloc(RIB_CAP * sizeof(rib_t));
    for (;;) {
        int64_t b, e2, i;
        pthread_mutex_lock(&j->mu);
        b = j->next; j->next += 16;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->n) break;
        e2 = b + 16 < j->n ? b + 16 : j->n;
        for (i = b; i < e2; i++) {
            int nr = 0;
            true_cost_one(j->c, &j->edges[i], &j->results[i], ribs, &nr);
            if (j->keep) {
                j->c->n_after[i] = nr;
                j->c->after[i] = (rib_t*)malloc((size_t)(nr > 0 ? nr : 1) * sizeof(rib_t));
                memcpy(j->c->after[i], ribs, (size_t)nr * sizeof(rib_t));
            }
        }
    }
    free(ribs);
    return NULL;
}

/* threads <= 0: all online cores; keep_ribbons: retain ribbons-after for oracle_get_ribbons_after */
int oracle_true_cost_batch_mt(oracle_ctx* c, int64_t n, const ppe_edge* edges, ppe_edge_result* results,
                              int threads, int keep_ribbons) {
    job_t j;
    int nt = threads, t;
    int64_t i;
    if (!c->have_cfg) { snprintf(c->err, sizeof c->err, "set_config first"); return PPE_ERR_STATE; }
    for (i = 0; i < n; i++)
        if (edges[i].ribbon_set < 0 || edges[i].ribbon_set >= c->n_sets) {
            snprintf(c->err, sizeof c->err, "edge %lld: unknown ribbon set %d", (long long)i, edges[i].ribbon_set);
            return PPE_ERR_INVALID;
        }
    free_after(c);
    if (keep_ribbons) {
        c->after = (rib_t**)calloc((size_t)(n > 0 ? n : 1), sizeof(rib_t*));
        c->n_after = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
        c->n_last = n;
    }
    if (nt <= 0) nt = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nt < 1) nt = 1;
    j.c = c; j.edges = edges; j.results = results; j.n = n; j.keep = keep_ribbons; j.next = 0;
    pthread_mutex_init(&j.mu, NULL);
    if (nt == 1) {
        worker(&j);
    } else {
        pthread_t* th = (pthread_t*)malloc((size_t)nt * sizeof(pthread_t));
        for (t = 0; t < nt; t++) pthread_create(&th[t], NULL, worker, &j);
        for (t = 0; t < nt; t++) pthread_join(th[t], NULL);
        free(th);
    }
    pthread_mutex_destroy(&j.mu);
    return PPE_OK;
}

int oracle_true_cost_batch(oracle_ctx* c, int64_t n, const ppe_edge* edges, ppe_edge_result* results) {
    return oracle_true_cost_batch_mt(c, n, edges, results, 1, 1);
}

int oracle_get_ribbons_after(oracle_ctx* c, int64_t i, double* xyxy, int cap) {
    int k, n;
    if (i < 0 || i >= c->n_last || !c->after) return PPE_ERR_INVALID;
    n = c->n_after[i];
    for (k = 0; k < n && k < cap; k++) memcpy(xyxy + 4 * k, &c->after[i][k], sizeof(rib_t));
    return n;
}

int oracle_max_threads(void) {
    int n = (int)sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : n;
}

/* ------------------------------------------------------------------------------------------- */
/* Direct access to the restated primitives, for the transcribed reference unit tests            */
/* (tests/test_golden_reference.py) and the host-side pinning of the engine's scalar helpers.    */
/* ------------------------------------------------------------------------------------------- */

/* RibbonManager::approximateDistanceUntilDone with MaxDistance (RibbonManager.cpp:28-36,234-248) */
double oracle_max_distance(oracle_ctx* c, int set, double x, double y) {
    const ribset_t* s = &c->sets[set];
    if (s->n == 0) return 0;
    return ribs_max_distance(s->r, s->n, x, y, c->cfg.ribbon_width);
}

/* RibbonManager::minDistanceFrom (RibbonManager.cpp:142-152) */
double oracle_min_distance_from(oracle_ctx* c, int set, double x, double y) {
    const ribset_t* s = &c->sets[set];
    return ribs_min_distance_from(s->r, s->n, x, y, c->cfg.ribbon_width);
}

/* RibbonManager::cover applied to a stored set (in place).  Returns the new ribbon count. */
int oracle_cover(oracle_ctx* c, int set, double x, double y, int strict) {
    ribset_t* s = &c->sets[set];
    rib_t* tmp = (rib_t*)malloc(RIB_CAP * sizeof(rib_t));
    int n = s->n;
    memcpy(tmp, s->r, (size_t)n * sizeof(rib_t));
    if (ribs_cover(tmp, &n, RIB_CAP, x, y, strict, c->cfg.ribbon_width)) { free(tmp); return -1; }
    free(s->r);
    s->r = (rib_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(rib_t));
    memcpy(s->r, tmp, (size_t)n * sizeof(rib_t));
    s->n = n;
    free(tmp);
    return n;
}

int oracle_get_ribbon_set(oracle_ctx* c, int set, double* xyxy, int cap) {
    const ribset_t* s = &c->sets[set];
    int k;
    for (k = 0; k < s->n && k < cap; k++) memcpy(xyxy + 4 * k, &s->r[k], sizeof(rib_t));
    return s->n;
}

/* Ribbon::split on a free-standing ribbon (Ribbon.cpp:9-17): rib = sx sy ex ey (updated in place),
 * piece receives the split-off part. */
void oracle_ribbon_split(oracle_ctx* c, double* rib, double x, double y, int strict, double* piece) {
    rib_t r, p;
    memcpy(&r, rib, sizeof r);
    p = rib_split(&r, x, y, strict, c->cfg.ribbon_width);
    memcpy(rib, &r, sizeof r);
    memcpy(piece, &p, sizeof p);
}

double oracle_collision_exists(oracle_ctx* c, double x, double y, double t, int strict) {
    return collision_exists(c, x, y, t, strict);
}

int oracle_is_blocked(oracle_ctx* c, double x, double y) { return map_blocked(c, x, y); }

/* DubinsWrapper::sample for a filled wrapper (DubinsWrapper.cpp:29-49,84-93); ok[i] = 0 where the
 * reference throws. */
void oracle_wrapper_sample(const double* qi, const double* param, double rho, int type, double w_start,
                           double w_speed, int n, const double* times, double* x, double* y, double* heading,
                           int32_t* ok) {
    wrapper_t w;
    int i;
    wrapper_init(&w);
    w.path.qi[0] = qi[0]; w.path.qi[1] = qi[1]; w.path.qi[2] = qi[2];
    w.path.param[0] = param[0]; w.path.param[1] = param[1]; w.path.param[2] = param[2];
    w.path.rho = rho; w.path.type = (DubinsPathType)type;
    w.speed = w_speed; w.start = w.ustart = w_start;
    w.end = w.start + dubins_path_length(&w.path) / w.speed;
    for (i = 0; i < n; i++) {
        double pose[4] = {0, 0, 0, 0};
        ok[i] = wrapper_sample(&w, pose, times[i]) ? 0 : 1;
        x[i] = pose[0]; y[i] = pose[1]; heading[i] = pose[2];
    }
}
