/* ORACLE / TEST INFRASTRUCTURE -- plain-C restatement of the ikd-Tree hot path (see ikd_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it; never the product. */
#ifndef IKD_ORACLE_H_
#define IKD_ORACLE_H_
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ikdo_tree ikdo_tree;

ikdo_tree* ikdo_create(float delete_param, float balance_param, float box_length);
void ikdo_destroy(ikdo_tree* t);
void ikdo_set_params(ikdo_tree* t, float delete_param, float balance_param, float box_length);
void ikdo_build(ikdo_tree* t, const float* xyz, long n);
int ikdo_knn(ikdo_tree* t, const float* q, int k, double max_dist, float* out_xyz, float* out_d);
int ikdo_knn_batch(ikdo_tree* t, const float* q, long nq, int k, double max_dist, float* out_xyz, float* out_d,
                   int* out_cnt, int nthreads);
long ikdo_box_search(ikdo_tree* t, const float* box6, float* out_xyz, long cap);
long ikdo_radius_search(ikdo_tree* t, const float* c, float r, float* out_xyz, long cap);
long ikdo_last_result(ikdo_tree* t, float* out_xyz, long cap);
int ikdo_add_points(ikdo_tree* t, const float* xyz, long n, int downsample_on);
void ikdo_delete_points(ikdo_tree* t, const float* xyz, long n);
int ikdo_delete_boxes(ikdo_tree* t, const float* boxes, long nb);
void ikdo_add_boxes(ikdo_tree* t, const float* boxes, long nb);
int ikdo_size(ikdo_tree* t);
int ikdo_validnum(ikdo_tree* t);
void ikdo_root_alpha(ikdo_tree* t, float* bal, float* del);
void ikdo_tree_range(ikdo_tree* t, float* box6);
long ikdo_flatten(ikdo_tree* t, float* out_xyz, long cap);
long ikdo_acquire_removed(ikdo_tree* t, float* out_xyz, long cap);
long ikdo_dump_tree(ikdo_tree* t, float* out, long cap);
int ikdo_max_depth(ikdo_tree* t);
double ikdo_mean_visits(ikdo_tree* t, const float* q, long nq, int k, double max_dist);
int ikdo_num_threads(void);
int ikdo_rebuild_count(ikdo_tree* t);
/* plane fit of the caller's next step; see ikd_oracle.c */
int ikdo_plane_fit(const float* nbr, int k, float thr, float* plane4);
void ikdo_plane_batch(const float* q, long nq, int k, const float* nbr, const float* sqd, const int* cnt,
                      float max_kth_sqdist, float thr, float* out_plane, float* out_resid, unsigned char* out_valid);

#ifdef __cplusplus
}
#endif
#endif
