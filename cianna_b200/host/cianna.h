/*
 * cianna.h - host-side C library of the B200-native CIANNA core (libcianna_host.so).
 *
 * Mirrors the operator interface of the reference's host layer for the hot path
 * (names, argument order/meaning and exit-on-error behaviour of src/prototypes.h:36-110:
 * init_network, create_dataset, conv_create, pool_create, norm_create, dense_create,
 * train_network, forward_testset, compute_error, save_network, load_network, set_frozen_layers),
 * so that upstream call sites (src/python_module.c, src/main.c) and the parity tests read the same
 * on both sides.  The implementation is new: every layer object drives the C-ABI of
 * include/cianna_b200.h; there is exactly one compute method, "C_CUDA", and no CPU fallback.
 *
 * Object model differences with upstream (internal only): activations live on the device in the
 * core's channels-last layout, layer->output / layer->delta_o are device pointers in that layout;
 * use cb_layer_export_* to read them back in the reference's [C][B][H*W] / [B][n+1] layouts.
 */
#ifndef CIANNA_HOST_H
#define CIANNA_HOST_H

#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include "cianna_b200.h"

#ifndef MAX_LAYERS_NB
#define MAX_LAYERS_NB 200
#endif
#ifndef MAX_NETWORKS_NB
#define MAX_NETWORKS_NB 10
#endif

/* same vocabulary / numeric values as upstream src/structs.h:32-43,69-70 (they appear in save files
 * and in the Python-level strings) */
enum layer_type_enum { CONV, POOL, DENSE, NORM, LRN };
enum activation_functions_enum { RELU, LOGISTIC, SOFTMAX, YOLO, LINEAR };
enum inference_modes_enum { AVG_MODEL, MC_MODEL };
enum batch_param_enum { OFF, SGD, FULL };
enum compute_method_enum { C_NAIV, C_BLAS, C_CUDA };
enum memory_localization_enum { NO_LOC, HOST, DEVICE };
enum pool_types_enum { MAX_pool, AVG_pool };
enum TC_comp_mode { FP32C_FP32A, TF32C_FP32A, FP16C_FP32A, FP16C_FP16A, BF16C_FP32A };

#define CB_DP_MAX_BUCKETS 8
typedef struct network network;
typedef struct layer layer;
typedef struct yolo_param yolo_param;

typedef struct Dataset {
	int size;              /* number of samples */
	int nb_batch;
	int localization;      /* NO_LOC / HOST / DEVICE */
	void **input;          /* [nb_batch] pinned host batches, compute dtype, [batch_size][input_dim+1] (bias slot last) */
	void **target;         /* [nb_batch] pinned host batches, compute dtype, [batch_size][output_dim] */
	void **input_device;   /* device-resident copies when dynamic_load == 0 */
	void **target_device;
	void *shuffle_ws;      /* device-side shuffle state of a resident set (duplicate batches, pointer tables, index) */
	int host_stale;        /* the resident copy was permuted on the device: the host batches follow on demand */
} Dataset;

struct layer {
	int type;
	int activation_type;
	int index;
	network *c_network;
	layer *previous;
	void *param;
	void *output;          /* device, internal layout [B][h][w][Cp] */
	void *delta_o;         /* device, same shape (NULL when inference_only) */
	int out_c, out_h, out_w;   /* out_h counts ROWS: depth x height for a 3-D map (activations [B][D][H][W][Cp]); only conv and
	                            * pool layers look inside, every other layer sees out_h * out_w pixels */
	int out_d;                 /* depth of the map (1 for 2-D networks); true height = out_h / out_d */
	int frozen;
	float bias_value;
	float dropout_rate;
	cb200_activ activ;
	void *activ_param;     /* yolo_param of a YOLO output layer (own copy, device tables attached), else NULL */
	cb200_dropout_desc drop; /* dropout_rate > 0.01: mask + activation pass run after the layer's own kernel (cb_dropout_*) */
	void (*forward)(layer *current);
	void (*backprop)(layer *current);
	int nb_params;
	float time_fwd, time_back;
};

typedef struct conv_param {
	int f_size[3], stride[3], padding[3], int_padding[3];
	int nb_filters, nb_area[3], prev_size[3], prev_depth;
	int flat_f_size;
	cb200_conv_desc desc;
	cb200_conv_weights w;
	int bias_grad_from_next;   /* the following norm layer's backward pass also produces this layer's grad_b */
	size_t grad_offset;    /* position of [grad | grad_b] in the network's gradient arena */
	size_t grad_len;
} conv_param;

typedef struct pool_param {
	int p_size[3], stride[3], padding[3], nb_area[3], prev_size[3];
	int nb_maps, prev_depth, pool_type, global;
	cb200_pool_desc desc;
	uint8_t *pool_map;     /* device */
	int fused_norm;        /* the preceding group-norm layer is evaluated inside this layer's kernels (see norm_param) */
} pool_param;

typedef struct norm_param {
	int group_size, set_off, nb_group, n_dim, dim_offset;
	cb200_norm_desc desc;
	float *gamma, *beta, *gamma_update, *beta_update;   /* device, FP32 [nb_group] */
	float *mean, *var, *d_gamma, *d_beta;               /* device, FP32 [batch][nb_group] */
	float *gsum;                                        /* device, FP32 [2][nb_group] inside the gradient arena */
	void *workspace;
	size_t grad_offset;
	/* set when the next layer is a 2x2 max-pool: the pair runs as cb200_norm_pool_forward / _backward, this layer's
	 * full-resolution output and delta are never materialised (layer->output == layer->delta_o == NULL) */
	layer *fused_pool;
	/* one-shot: the convolution in front has just left this pass's (sum, sum of squares) in `workspace`
	 * (cb200_conv_forward_stats), the forward call skips its statistics launch */
	int stats_ready;
} norm_param;

/* local response normalisation across channels (src/structs.h lrn_param, src/lrn_layer.c) */
typedef struct lrn_param {
	int range, n_dim, dim_offset;
	float k, alpha, beta;
	cb200_lrn_desc desc;
	float *local_scale;    /* device, FP32, one value per activation (training only) */
} lrn_param;

typedef struct dense_param {
	int in_size;           /* inputs incl. the bias node */
	int nb_neurons;
	int prev_c, prev_h, prev_w;
	cb200_conv_desc desc;  /* a dense layer runs as a convolution whose filter covers the whole input map */
	cb200_conv_weights w;
	size_t grad_offset, grad_len;
} dense_param;

/* YOLO output-layer set-up, same fields and defaults as upstream's yolo_param (src/structs.h:527-575); the
 * association scratch tables of upstream are replaced by one device workspace (include/cianna_b200.h) */
struct yolo_param {
	int no_override, raw_output;
	int nb_box, nb_class, nb_param, max_nb_obj_per_image, fit_dim;
	int IoU_type, prior_dist_type;
	float *prior_size;            /* host [nb_box][3] */
	float *noobj_prob_prior;      /* host [nb_box] */
	int class_softmax, diff_flag, error_type;
	int strict_box_size_association, rand_startup;
	float rand_prob_best_box_assoc, rand_prob, min_prior_forced_scaling;
	float scale_tab[6];
	float slopes_and_maxes_tab[6][3];
	float *param_ind_scale;       /* host [nb_param] */
	float IoU_limits[8];
	int fit_parts[6];
	int cell_size[3];
	/* device side (filled by set_yolo_activ on the layer's copy) */
	cb200_yolo_desc desc;
	float *dev_tables;            /* prior_size | noobj_prob_prior | param_ind_scale */
	float *workspace;
	float *parts_dev, *parts_host;       /* [batch][6] loss split */
	float *monitor_dev, *monitor_host;   /* [batch][cells][nb_box][2] */
	int *box_state_dev;                  /* [batch][cells][nb_box] lock state written by the association pass */
	unsigned long long seed, step;
};

struct network {
	layer *net_layers[MAX_LAYERS_NB];
	int id;
	int compute_method;
	int inference_only;
	int nb_layers;
	float input_bias;
	float learning_rate, momentum, decay, weight_decay;
	Dataset train, test, valid;
	Dataset train_buf, test_buf, valid_buf;
	int in_dims[4];
	size_t input_dim;
	int output_dim;
	int out_size;
	int batch_size;
	int batch_param;
	int iter;
	int is_inference;
	int inference_drop_mode;
	int no_error;
	int perf_eval;
	long long int total_nb_param;
	long long int memory_footprint;
	int adv_size;
	int length;
	float TC_scale_factor;
	int dynamic_load;
	int use_cuda_TC;       /* enum TC_comp_mode */
	int dtype;             /* cb200_dtype derived from use_cuda_TC */

	/* device-side state of one step */
	void *input_raw;       /* current batch in dataset layout (device) */
	void *input;           /* current batch in internal layout (device) */
	void *target;          /* current target batch (device, compute dtype) */
	float *loss_dev;       /* FP32 [batch_size] per-sample loss */
	float *loss_host;      /* pinned */
	float *hyper_dev;      /* CB200_HYPER_LEN floats */
	/* the optimizer of every layer but the first starts while the first layer's weight gradient - the last kernel of the
	 * backward sweep, alone on the side stream - is still running: `upper_ev` is recorded on the side stream before that
	 * kernel is enqueued (all other weight gradients are ahead of it), the compute stream waits for the event only */
	void *upper_ev;
	int upper_marked;
	void *update_plan;     /* cb200_update_plan of the layers below (the optimizer sweep in three launches), or NULL */
	unsigned update_plan_sig;   /* which layers were frozen / how the norm sums arrive when the plan was built */
	unsigned char in_update_plan[MAX_LAYERS_NB];
	float *grad_arena;     /* all raw gradients, contiguous (all-reduced in data-parallel runs) */
	size_t grad_arena_len;
	int training_ready;
	int dp_world;          /* data-parallel world size (1 = single GPU) */
	/* gradient exchange buckets (cb_dp_plan): contiguous arena slices, each all-reduced in ONE call as soon as its
	 * lowest layer's weight gradient is enqueued (the backward sweep runs from the last layer down) */
	int dp_nb_bucket;
	size_t dp_bucket_begin[CB_DP_MAX_BUCKETS], dp_bucket_len[CB_DP_MAX_BUCKETS];
	int dp_bucket_trigger[CB_DP_MAX_BUCKETS];      /* index of the layer whose backward pass issues the bucket */
	float last_batch_loss;
	double last_epoch_loss;
	float last_items_per_s;
	double last_accuracy;  /* of the last compute_error run with a confusion matrix (fraction of correct argmax) */
	void *out_host;        /* pinned staging of the last layer's output (inference read-back) */
	/* double-buffered host->device staging of dynamic_load batches on a copy stream (overlaps the previous step) */
	void *copy_stream;
	/* weight-gradient kernels run on their own stream: the rest of the backward sweep does not depend on them, and the
	 * bandwidth-bound group-norm / pool kernels of the layers below fill the SMs next to the tensor-core bound wgrad CTAs */
	void *wgrad_stream;
	void *wgrad_stream_off;    /* parked here while cb_set_wgrad_overlap(net, 0) is in effect */
	void *stage_in[2], *stage_tg[2];
	const void *staged_src[2];   /* host batch currently (being) copied into each slot */
	int stage_slot;
	const cb200_conv_desc *patch_desc;   /* first conv layer when it consumes patch rows (few input channels), else NULL */
	yolo_param *y_param;   /* network-level YOLO set-up (set_yolo_params), copied into the YOLO layer at creation */
	/* per-layer timing table of perf_eval (src/auxil.c:698-870 upstream): sampled on the first mini-batch of every epoch
	 * with events between the layers (3 x (nb_layers + 1): forward, backward, optimizer), no synchronisation inside */
	void **perf_ev;
	int perf_sample;
	double *fwd_perf, *back_perf;   /* accumulated microseconds per layer */
	int perf_n;
	unsigned long long drop_seed;   /* dropout masks are a function of (seed, layer, draw, position): see cb200_dropout_desc */
	unsigned long long drop_draw;   /* counts the forward passes that drew masks */
};

extern network *networks[MAX_NETWORKS_NB];
extern int nb_networks;
extern int is_init;

/* ---- upstream-compatible API (src/prototypes.h) ---- */
void init_network(int network_number, int u_input_dim[4], int u_output_dim, float in_bias, int u_batch_size,
	const char *compute_method_string, int u_dynamic_load, const char *cuda_TC_string, int inference_only, int no_logo, int adv_size);
Dataset create_dataset(network *net, int nb_elem);
void free_dataset(Dataset *data);
int nb_area_comp(int size, int f_size, int padding, int int_padding, int stride);
int conv_create(network *net, layer *previous, int *f_size, int nb_filters, int *stride, int *padding,
	int *int_padding, int *in_shape, const char *activation, float *bias, float drop_rate,
	const char *init_fct, float init_scaling, FILE *f_load, int f_bin);
int pool_create(network *net, layer *previous, int *pool_size, int *stride, int *padding,
	const char *char_pool_type, const char *activation, int global, float drop_rate);
int norm_create(network *net, layer *previous, const char *norm_type, const char *activation, int group_size, int set_off, FILE *f_load, int f_bin);
int dense_create(network *net, layer *previous, int nb_neurons, const char *activation, float *bias,
	float drop_rate, int strict_size, const char *init_fct, float init_scaling, FILE *f_load, int f_bin);
void conv_save(FILE *f, layer *current, int f_bin);
void conv_load(network *net, FILE *f, int f_bin);
void pool_save(FILE *f, layer *current, int f_bin);
void pool_load(network *net, FILE *f, int f_bin);
void norm_save(FILE *f, layer *current, int f_bin);
void norm_load(network *net, FILE *f, int f_bin);
int lrn_create(network *net, layer *previous, const char *activation, int range, float k, float alpha, float beta, FILE *f_load, int f_bin);
void lrn_save(FILE *f, layer *current, int f_bin);
void lrn_load(network *net, FILE *f, int f_bin);
int cb_lrn_range(layer *current);
/* dropout of a layer's output (rate > 0.01 as upstream): set-up at creation, mask/scale + activation after the layer's
 * forward kernel, mask on the layer's delta before its backward kernels (layers.c) */
void cb_dropout_setup(layer *current);
void cb_dropout_forward(layer *current);
void cb_dropout_backward(layer *current);
void cb_set_dropout_seed(network *net, unsigned long long seed);
void cb_set_inference_drop_mode(network *net, int mode);
void cb_layer_export_dropout_mask(network *net, int l, float *dst);
void dense_save(FILE *f, layer *current, int f_bin);
void dense_load(network *net, FILE *f, int f_bin);
void print_architecture_tex(network *net, const char *path, const char *file_name, int l_size, int l_in_size,
	int l_f_size, int l_out_size, int l_stride, int l_padding, int l_in_padding, int l_activation, int l_bias,
	int l_dropout, int l_param_count);
void save_network(network *net, const char *filename, int f_bin);
void load_network(network *net, const char *filename, int epoch, int nb_layers, int f_bin);
void set_frozen_layers(network *net, int *tab, int dim);
void train_network(network *net, int nb_epochs, int control_interv, float u_begin_learning_rate, float u_end_learning_rate, float u_momentum,
	float u_decay, float u_weight_decay, int show_confmat, int save_net, int save_bin, int shuffle_gpu, int shuffle_every, float c_TC_scale_factor, int silent);
void forward_testset(network *net, int saving, int repeat, int drop_mode, int silent);
void compute_error(network *net, Dataset data, int saving, int confusion_matrix, int repeat, int silent);
void perf_eval_display(network *net);
int cb_perf_eval_read(network *net, double *fwd, double *back);

/* activation helpers (src/activ_functions.c:260-374, 580-610) */
void load_activ_param(layer *current, const char *activ);
void set_activ_defaults(layer *current, const char *activ);
void print_string_activ_param(layer *current, char *activ);
void print_activ_param(FILE *f, layer *current, int f_bin);
/* YOLO output layer (src/activ_functions.c:970-1477) */
int set_yolo_params(network *net, size_t nb_box, int nb_class, int nb_param, int max_nb_obj_per_image, const char *IoU_type_char,
	const char *prior_dist_type_char, float *prior_size, float *yolo_noobj_prob_prior, int fit_dim,
	int strict_box_size, int rand_startup, float rand_prob_best_box_assoc, float rand_prob, float min_prior_forced_scaling, float *scale_tab,
	float **slopes_and_maxes_tab, float *param_ind_scale, float *IoU_limits, int *fit_parts, int class_softmax,
	int diff_flag, const char *error_type, int no_override, int raw_output);
void set_yolo_activ(layer *current);
/* initialisers (src/initializers.c) */
void init_weights(float *tab, int dim_in, int dim_out, const char *init_fct, float init_scaling);

/* ---- additions of this implementation ---- */
/* copy one sample (FP32) into a dataset batch slot, converting to the compute dtype like upstream's
 * Dataset.cont_copy (src/cuda/cuda_main.cu:355-371) */
void dataset_set_sample(network *net, Dataset *data, int index, const float *input, const float *target);
void dataset_upload(network *net, Dataset *data);   /* dynamic_load == 0: make device-resident copies */
void shuffle_dataset(network *net, Dataset *data);  /* what train_network does every shuffle_every epochs */
void shuffle_dataset_device(network *net, Dataset *data);   /* ... with shuffle_gpu = 1 on a device-resident set */
void dataset_host_refresh(network *net, Dataset *data);     /* host batches <- resident copy after a device shuffle */
void cb_dataset_read_row(network *net, Dataset *data, int index, int which, int from_device, void *dst);
void cb_net_io_dims(network *net, long long *out3);   /* input_dim, output_dim, cb200 dtype */
/* data-parallel set-up: call on every rank after init_network, before training */
void cb_dp_unique_id(void *id128);
void cb_dp_init(network *net, const void *id128, int rank, int world);
/* Pure planning arithmetic (no device needed; also driven by tests/test_dp_gloo.py): lays out the gradient arena -
 * `head` floats of group-norm sums first, then the n weight-gradient slices of len[i] floats (each rounded up to 64)
 * in layer order - and cuts it into at most CB_DP_MAX_BUCKETS contiguous buckets for the exchange.  offset[i] receives
 * slice i's position; bucket b is [begin[b], begin[b] + blen[b]) and is complete once slice first[b] (its lowest) is
 * written.  Buckets are listed in the order they complete (last layers first).  Returns the number of buckets. */
int cb_dp_plan(int n, const size_t *len, size_t head, size_t *offset, size_t *begin, size_t *blen, int *first);
/* called by every conv / dense layer's backward pass once its weight gradient is enqueued */
void cb_dp_layer_done(network *net, layer *current);
/* one mini-batch from explicit host arrays (FP32, dataset layout [B][input_dim+1] and [B][output_dim]) */
void cb_load_batch(network *net, const float *input, const float *target);
void cb_load_batch_typed(network *net, const void *input_typed, const void *target_typed);
void cb_forward(network *net, int length, int is_inference);
void cb_backward(network *net, float lr, float momentum, float weight_decay);   /* delta + backprop + update */
float cb_batch_loss(network *net);                                               /* sum over samples / length */
void cb_train_step(network *net, float lr, float momentum, float weight_decay); /* forward + backward on the loaded batch */
void cb_sync(void);
void cb_set_update_plan(int on);   /* optimizer sweep in three launches (csrc/update_plan.cu) or layer by layer */
/* read-back in the reference's layouts; dst sized by the caller */
void cb_layer_export_output(network *net, int l, float *dst);
void cb_layer_export_delta(network *net, int l, float *dst);
void cb_layer_export_pool_map(network *net, int l, int *dst);
size_t cb_layer_weight_count(network *net, int l);
void cb_layer_get_weights(network *net, int l, float *dst);   /* conv/dense: FP32 master (reference layout); norm: gamma then beta */
void cb_layer_set_weights(network *net, int l, const float *src);
void cb_layer_get_moment(network *net, int l, float *dst);
void cb_layer_get_norm_stats(network *net, int l, float *mean, float *var, float *d_gamma, float *d_beta);
void cb_layer_shape(network *net, int l, int *out4);         /* c, h, w, type */
const char *cb_layer_conv_impl(network *net, int l);          /* which kernel family ran last ("tcgen05"/"simt") */

/* accessors for language bindings */
network *cb_get_network(int id);
int cb_net_nb_layers(network *net);
int cb_nb_networks(void);
layer *cb_net_layer(network *net, int idx);
int cb_net_batch_size(network *net);
float cb_net_last_items_per_s(network *net);
double cb_net_last_epoch_loss(network *net);
double cb_net_last_accuracy(network *net);
void cb_net_set_no_error(network *net, int v);
Dataset *cb_net_dataset(network *net, const char *name);
void cb_set_dataset(network *net, const char *name, int size, const float *input, const float *target);
void cb_swap_data_buffers(network *net, const char *name);
void cb_net_in_dims(network *net, int *out4);
void cb_set_TC_scale_factor(network *net, float v);
void cb_train_steps(network *net, int nsteps, float lr, float momentum, float weight_decay, int resident, int sync_each_step);
void cb_forward_steps(network *net, int nsteps, int resident, int sync_each_step);
/* YOLO read-backs for bindings / parity tests: box_state int32 [B][cells][nb_box] of the last cb_backward,
 * loss split [6] and monitor [B][cells][nb_box][2] of the last cb_batch_loss, decoded boxes [C][B][cells] */
void cb_yolo_set_seed(network *net, unsigned long long seed);
void cb_yolo_box_state(network *net, int *dst);
void cb_yolo_loss_parts(network *net, float *parts6, float *monitor);
void cb_yolo_export_boxes(network *net, float *dst);
void cb_net_set_iter(network *net, int iter, int train_size);
/* group-norm + max-pool fusion for the layers created from now on (default on; CB200_NO_FUSION=1 in the environment turns it off) */
void cb_set_fusion(int on);
/* 0: weight gradients back on the compute stream (one kernel at a time: what per-kernel event timing needs), 1: overlap */
void cb_set_wgrad_overlap(network *net, int on);

#define CB_CHECK(call) do { int rc__ = (call); if (rc__ != 0) { \
	printf("\nERROR: %s failed (%d): %s\n", #call, rc__, cb200_last_error()); exit(EXIT_FAILURE); } } while (0)

#endif
