/*
 * cianna_bridge.h - plain-argument entry points of libcianna_host.so for a back-end shim that lives INSIDE another
 * host program whose own symbols and struct names collide with cianna.h: upstream CIANNA itself.
 *
 * cianna_b200/shim/cuda_b200_shim.c implements the `cuda_*` symbols upstream's host code binds
 * (src/prototypes.h:217-295) and is compiled against upstream's structs.h; it cannot include cianna.h (same struct
 * names: network, layer, conv_param ...) nor call init_network / conv_create by name (upstream defines those too), so
 * everything it needs from the host library goes through the `cbb_*` functions below: ints, floats, raw pointers.
 * A network is addressed by its upstream id (networks[id] on both sides), a layer by its index.
 *
 * The mirror: for every upstream layer that is converted to the device (cuda_convert_X_layer) the shim creates the same
 * layer in the host library's network with upstream's geometry and upstream's initial weights; upstream's
 * layer->forward / layer->backprop then run the mirrored layer's kernels.
 */
#ifndef CIANNA_BRIDGE_H
#define CIANNA_BRIDGE_H

#include <stddef.h>

/* tc_mode: upstream enum TC_comp_mode (0 FP32C_FP32A, 1 TF32C_FP32A, 2 FP16C_FP32A, 3 FP16C_FP16A, 4 BF16C_FP32A) */
void cbb_init(int net_id, const int in_dims[4], int out_dim, float in_bias, int batch_size, int dynamic_load, int tc_mode,
	int inference_only, int adv_size);
int  cbb_nb_layers(int net_id);
int  cbb_dtype_size(int net_id);

/* layer creation; prev = index of the previous layer or -1; weights in upstream's host layouts, row_stride in floats
 * (conv: [nb_filters][flat_f_size + TC_padding]; dense: [in_size][nb_neurons + 1]); returns the layer index */
int  cbb_conv(int net_id, int prev, const int f_size[3], int nb_filters, const int stride[3], const int padding[3],
	const int int_padding[3], const char *activation, float bias, float drop_rate, const float *weights, int row_stride);
int  cbb_pool(int net_id, int prev, const int p_size[3], const int stride[3], const int padding[3], int is_avg,
	const char *activation, int global, float drop_rate);
int  cbb_norm(int net_id, int prev, const char *activation, int group_size, int set_off, const float *gamma, const float *beta);
int  cbb_lrn(int net_id, int prev, const char *activation, int range, float k, float alpha, float beta);
int  cbb_dense(int net_id, int prev, int nb_neurons, const char *activation, float bias, float drop_rate, const float *weights);
/* YOLO set-up with upstream's already-resolved numeric fields (src/structs.h:527-575); before the YOLO conv layer */
void cbb_set_yolo(int net_id, int nb_box, int nb_class, int nb_param, int max_nb_obj_per_image, int IoU_type,
	int prior_dist_type, const float *prior_size, const float *noobj_prob_prior, int fit_dim, int strict_box_size,
	int rand_startup, float rand_prob_best_box_assoc, float rand_prob, float min_prior_forced_scaling,
	const float *scale_tab6, const float *slopes_and_maxes_6x3, const float *param_ind_scale, const float *IoU_limits8,
	const int *fit_parts6, int class_softmax, int diff_flag, int error_type, int no_override, int raw_output);
void cbb_layer_shape(int net_id, int l, int *c_h_w);

/* device pointers upstream's structs keep (all FP32): conv / dense master weights in upstream's file layout WITHOUT the
 * TC padding columns, norm mean / var / d_gamma / d_beta [batch][nb_group], YOLO IoU monitor */
float *cbb_master(int net_id, int l);
float *cbb_moment(int net_id, int l);
float *cbb_norm_table(int net_id, int l, int what);   /* 0 mean, 1 var, 2 d_gamma, 3 d_beta, 4 gamma, 5 beta */
float *cbb_yolo_monitor(int net_id);
/* after a write into master weights from outside (cuda_put_table_FP32 / cuda_set_mem_value): rebuild the 16-bit operands */
void cbb_weights_changed(int net_id, int l);
void cbb_norm_set(int net_id, int l, const float *gamma, const float *beta);
void cbb_norm_get(int net_id, int l, float *gamma, float *beta);   /* synchronous */
void cbb_norm_get_async(int net_id, int l, float *gamma, float *beta);   /* queued on the compute stream; cbb_stream_sync() */
void cbb_stream_sync(void);
/* the layer's activation / error tensor in the core's own layout (a key for the caller's bookkeeping, not to be read
 * directly: cbb_export_act converts); NULL when the layer is evaluated inside its neighbour's kernel and keeps none */
void *cbb_act_ptr(int net_id, int l, int want_delta);

/* one mini-batch, layer by layer as upstream's loops call them (src/auxil.c:1827-1849) */
void cbb_forward_layer(int net_id, int l, const void *input_dev, int length, int is_inference, int mc_model);
void cbb_deriv_output_error(int net_id, const void *target_dev, float TC_scale_factor, int iter, int train_size);
void cbb_backprop_layer(int net_id, int l, float lr, float momentum, float weight_decay, int frozen);
/* per-element loss in upstream's layout of the last layer's output (FP32 device table, zeroed by the caller) */
void cbb_output_error(int net_id, const void *target_dev, float *err_dev, size_t err_elems);
/* activation tensors in upstream's layouts ([C][B][H*W] / dense [B][n+1]), FP32, to host (synchronous) */
void cbb_export_act(int net_id, int l, int want_delta, float *dst_host);
void cbb_sync(void);

#endif
