"""The reference's shipped pre-training hyper-parameters, as Python dicts (so that bench.py / smoke() do not need
a YAML file on the GPU box).  Values are those of configs_clean/pre-train_QM9.yml:50-102 (identical model blocks in
configs_clean/pre-train_QMugs.yml and the shipped checkpoint's train_arguments.yaml)."""

PRETRAIN_QM9_MODEL_PARAMETERS = dict(
    target_dim=256, hidden_dim=200, mid_batch_norm=True, last_batch_norm=True, readout_batchnorm=True,
    batch_norm_momentum=0.93, readout_hidden_dim=200, readout_layers=2, dropout=0.0, propagation_depth=7,
    aggregators=["mean", "max", "min", "std"], scalers=["identity", "amplification", "attenuation"],
    readout_aggregators=["min", "max", "mean"], pretrans_layers=2, posttrans_layers=1, residual=True)

PRETRAIN_QM9_MODEL3D_PARAMETERS = dict(
    target_dim=256, hidden_dim=20, hidden_edge_dim=20, node_wise_output_layers=0, message_net_layers=1,
    update_net_layers=1, reduce_func="mean", fourier_encodings=4, propagation_depth=1, dropout=0.0, batch_norm=True,
    readout_batchnorm=True, batch_norm_momentum=0.93, readout_hidden_dim=20, readout_layers=1,
    readout_aggregators=["min", "max", "mean"])

PRETRAIN_QM9 = dict(loss_func="NTXent", loss_params={"tau": 0.1}, optimizer="Adam", optimizer_params={"lr": 8.0e-5},
                    batch_size=500, model_type="PNA", model3d_type="Net3D")

# configs_clean/tune_QM9_homo.yml:10-74 (fine-tuning on the QM9 'homo' target from a pre-trained checkpoint)
TUNE_QM9_HOMO_MODEL_PARAMETERS = dict(
    target_dim=1, hidden_dim=200, mid_batch_norm=True, last_batch_norm=True, readout_batchnorm=True,
    batch_norm_momentum=0.1, readout_hidden_dim=200, readout_layers=2, dropout=0.0, propagation_depth=7,
    aggregators=["mean", "max", "min", "std"], scalers=["identity", "amplification", "attenuation"],
    readout_aggregators=["min", "max", "mean", "sum"], pretrans_layers=2, posttrans_layers=1, residual=True)

TUNE_QM9_HOMO = dict(loss_func="L1Loss", optimizer="Adam", optimizer_params={"lr": 7.0e-5, "weight_decay": 1.0e-11},
                     batch_size=128, model_type="PNA", transfer_layers=["gnn"], exclude_from_transfer=["batch_norm"])

# The tower PNA in the shape BASELINE.json configs[1] words ("hidden 200, 4 towers, 4 layers"): models/pna_original.py
# with the contrastive settings of the reference's tower configs, inputs divided between the towers
PNA_ORIGINAL_H200_T4_MODEL_PARAMETERS = dict(
    target_dim=256, hidden_dim=200, last_layer_dim=200, mid_batch_norm=True, last_batch_norm=True, graph_norm=False,
    readout_batchnorm=True, edge_hidden_dim=200, readout_hidden_dim=100, readout_layers=2, dropout=0.0,
    in_feat_dropout=0.0, propagation_depth=4, towers=4, divide_input_first=True, divide_input_last=True,
    aggregators=["mean", "max", "min", "std"], scalers=["identity", "amplification", "attenuation"],
    readout_aggregators=["mean", "max", "min", "sum"], pretrans_layers=1, posttrans_layers=1, residual=True)
