from volume_segmantics_b200.host.predict_cli import create_output_path, main  # noqa: F401

if __name__ == "__main__":
    main()
